"""Shared test helpers: workloads of the BASELINE configs and the parity comparison (SURVEY.md 8c tolerances)."""
from __future__ import annotations

import numpy as np

from live_ekf_slam_b200 import Params
from live_ekf_slam_b200 import workload as wl

REL_TOL = 1e-9          # per-step |delta| <= 1e-9 * max(1, |.|_max)   (north star)
FINAL_TOL = 1e-6        # final trajectories: 1e-6 m / 1e-6 rad        (north star)


def config1(seed=0, steps=1000):
    """single EKF run, random 20-landmark map, TSP trajectory, known IDs (filter_demo_live defaults)."""
    p = Params(filter="ekf_slam")
    rng = np.random.default_rng(seed)
    lm = wl.random_map(20, p.map_bound, p.map_min_landmark_separation, rng)
    fwd, ang = wl.tsp_trajectory(lm, p, rng, steps)
    return p, lm, fwd, ang


def config2(seed=0, steps=1000, filt="ekf_slam"):
    """Monte-Carlo batch on the 5x10 grid map with one shared TSP command trajectory."""
    p = Params(filter=filt)
    rng = np.random.default_rng(seed)
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, rng, steps)
    return p, lm, fwd, ang


def oracle_params(oc, p: Params):
    return oc.make_params(v_d=p.v_d, v_th=p.v_th, w_r=p.w_r, w_b=p.w_b, V_00=p.V_00, V_11=p.V_11, W_00=p.W_00,
                          W_11=p.W_11, landmark_id_is_known=int(p.landmark_id_is_known),
                          min_landmark_separation=p.min_landmark_separation, compat_noise_bug=int(p.compat_noise_bug),
                          d_max=p.d_max, th_max=p.th_max, range_max=p.range_max, fov_min=p.fov_min, fov_max=p.fov_max)


def normwise(a, b):
    """max |a-b| / max(1, max|b|)"""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def oracle_meas_stream(oc, op, lm, fwd, ang, seed, instance):
    """the simulator's messages for one vehicle: list of float32 [k,3] + truth trace"""
    truth = np.zeros(3)
    out, tr = [], []
    for t in range(len(fwd)):
        out.append(oc.sim_step(op, truth, fwd[t], ang[t], lm, seed, instance, t))
        tr.append(truth.copy())
    return out, np.asarray(tr)


#: non-default noise settings: the float adds `d_d + v_d` (ekf.cpp:57-58), `u_d + v_d` (ukf.cpp:129-131), the
#: all-float innovation `r - dist - w_r` (ekf.cpp:130-131), the sensing-model offsets (ukf.cpp:144-145) and the corrected-noise
#: branch of readCommonParams (filter.h:105-121 without the V/W mix-up) are only exercised away from the yaml defaults.
PARAM_VARIANTS = {
    "noise_means": dict(v_d=0.003, v_th=-0.002, w_r=0.004, w_b=-0.003),
    "no_noise_bug": dict(compat_noise_bug=False),
    "no_noise_bug_means_covs": dict(compat_noise_bug=False, v_d=-0.002, v_th=0.0015, w_r=-0.005, w_b=0.002,
                                    V_00=0.02, V_11=0.002, W_00=0.02, W_11=0.005),
}


def variant_params(name, filt="ekf_slam", known=True):
    p = Params(filter=filt)
    for k, v in PARAM_VARIANTS[name].items():
        setattr(p, k, v)
    p.landmark_id_is_known = known
    return p


def variant_workload(name, filt="ekf_slam", known=True, seed=0, steps=200):
    """5x10 grid + shared TSP trajectory (config 2 shape) under a non-default noise setting"""
    p = variant_params(name, filt, known)
    rng = np.random.default_rng(seed)
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, rng, steps)
    return p, lm, fwd, ang


def device_message_stream(shim, p, lm, fwd, ang, B, seed, offset, max_meas=8, max_lm=50):
    """The messages the ON-DEVICE simulator emits for B vehicles, recorded step by step from a twin handle (same Philox
    streams as any other handle with the same seed / offset): list over t of (meas [B][max_meas][3], n [B]) plus the truth
    trace [T][B][3].  Feeding the oracle the device's own messages keeps free-running comparisons at the 1e-9 bar: the
    device libm and glibc differ in the last float32 bit of r / b on a few of every several thousand messages."""
    tw = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, max_lm, max_meas)
    tsim = shim.Simulator(tw, lm, seed=seed, instance_offset=offset)
    msgs, truth = [], []
    for t in range(len(fwd)):
        tsim.step(fwd[t], ang[t], t)
        m, n = tsim.meas()
        msgs.append((m.copy(), n.copy()))
        truth.append(tsim.truth().copy())
    tsim.close()
    tw.close()
    return msgs, np.asarray(truth)
