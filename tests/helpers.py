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
