"""generate_landmarks (sim_node.py:155-206): the CPU restatement (oracle_make_map) against numpy / a pure-Python transcription,
and (GPU) the device generator slam_sim_make_maps against the restatement bit for bit, then a Monte-Carlo sweep over
per-instance random maps against oracle runs on those maps."""
import math

import numpy as np
import pytest

from tests import helpers as H


def test_grid_map_matches_numpy_arange(oracle):
    from live_ekf_slam_b200 import workload as wl
    for bound, step in ((10.0, 4.0), (10.0, 3.0), (6.5, 1.3)):
        p = H.Params()
        p.map_bound, p.map_grid_step = bound, step
        ref = wl.grid_map(p)                                        # np.arange(-bound + step / 2, bound, step) squared
        got = oracle.make_map("grid", 0, bound, step, 0.05, 0, 0)
        np.testing.assert_array_equal(got, ref)


def test_random_map_restatement(oracle):
    seed, inst, n, bound, sep = 99, 7, 40, 10.0, 0.8
    got = oracle.make_map("random", n, bound, 4.0, sep, seed, inst)
    # pure-Python transcription of sim_node.py:177-188 on the same Philox draws
    pts, a = [], 0
    while len(pts) < n:
        r = oracle.philox(inst, a, 0, 2, seed & 0xffffffff, seed >> 32)
        a += 1
        pos = (2 * bound * oracle.uniform(r[0], r[1]) - bound, 2 * bound * oracle.uniform(r[2], r[3]) - bound)
        if any(((q[0] - pos[0]) ** 2 + (q[1] - pos[1]) ** 2) ** (1 / 2) < sep for q in pts):
            continue
        pts.append(pos)
    np.testing.assert_array_equal(got, np.array(pts))
    assert a > n                                                    # some candidates were rejected: the separation test is live
    d = np.hypot(got[:, None, 0] - got[None, :, 0], got[:, None, 1] - got[None, :, 1]) + np.eye(n) * 1e9
    assert d.min() >= sep and np.abs(got).max() < bound
    assert not np.array_equal(got, oracle.make_map("random", n, bound, 4.0, sep, seed, inst + 1))
    with pytest.raises(ValueError):
        oracle.make_map("hexagonal", n, bound, 4.0, sep, seed, inst)
    with pytest.raises(ValueError):
        oracle.make_map("random", 50, 1.0, 4.0, 5.0, seed, inst)   # cannot be completed


@pytest.mark.gpu
def test_device_maps_and_sweep_over_random_maps(oracle):
    from live_ekf_slam_b200 import shim
    shim.load()
    p = H.Params(filter="ekf_slam")
    op = H.oracle_params(oracle, p)
    B, seed, off, n_lm = 10, 2024, 3, 20
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 32, 8)
    fb.init(*p.init_pose)
    sim = shim.Simulator(fb, np.zeros((1, 2)), seed=seed, instance_offset=off)
    # grid: params.yaml defaults (bound 10, step 4) -> 25 landmarks, identical for every vehicle
    assert sim.make_maps("grid", 0, p.map_bound, p.map_grid_step) == 25
    np.testing.assert_array_equal(sim.get_map(0), oracle.make_map("grid", 0, p.map_bound, p.map_grid_step, 0.05, seed, off))
    np.testing.assert_array_equal(sim.get_map(B - 1), sim.get_map(0))
    # random: params.yaml defaults (20 landmarks, min separation 0.05 is too loose to ever reject: use 1.5 as well)
    for sep in (p.map_min_landmark_separation, 1.5):
        assert sim.make_maps("random", n_lm, p.map_bound, p.map_grid_step, sep) == n_lm
        for i in range(B):
            np.testing.assert_array_equal(sim.get_map(i), oracle.make_map("random", n_lm, p.map_bound, p.map_grid_step, sep, seed, off + i))
    assert not np.array_equal(sim.get_map(0), sim.get_map(1))
    with pytest.raises(shim.SlamError):
        sim.make_maps("hexagonal", n_lm, p.map_bound)
    with pytest.raises(shim.SlamError):
        sim.make_maps("random", 50, 1.0, 4.0, 5.0)
    # a sweep on the per-instance maps (trajectories generated on the device from each vehicle's own map) against oracle runs
    import torch
    T = 200
    d_fwd = torch.zeros((T, B), dtype=torch.float32, device="cuda")
    d_ang = torch.zeros((T, B), dtype=torch.float32, device="cuda")
    sim.make_trajectories(p.landmark_noise, p.visitation_threshold, p.map_bound, p.init_pose, T, d_fwd, d_ang)
    sim.run_device(d_fwd, d_ang, 1, T, 0)
    fb.synchronize()
    fwd, ang = d_fwd.cpu().numpy(), d_ang.cpu().numpy()
    for i in (0, 4, 9):
        lm = oracle.make_map("random", n_lm, p.map_bound, p.map_grid_step, 1.5, seed, off + i)
        rf, ra = oracle.tsp_trajectory(op, lm, p.landmark_noise, p.visitation_threshold, p.map_bound, p.init_pose, T, seed, off + i)
        assert np.abs(fwd[:, i] - rf).max() <= 1e-7 and np.abs(ang[:, i] - ra).max() <= 1e-7
        # the oracle filter on the device's own messages (the device generator may flip a last float32 bit, test_gpu_sim_parity)
        tw = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 32, 8)
        tsim = shim.Simulator(tw, lm, seed=seed, instance_offset=off + i)
        of = oracle.OracleFilter(oracle.EKF_SLAM, op, 32)
        of.init(*p.init_pose)
        for t in range(T):
            tsim.step(fwd[t, i], ang[t, i], t)
            m, n = tsim.meas()
            of.update(fwd[t, i], ang[t, i], m[0, : n[0]], oracle.STRUCTURED)
        assert fb.num_landmarks(i) == of.M and list(fb.landmark_ids(i)) == list(of.landmark_ids()) and of.M >= 5
        assert H.normwise(fb.state(i), of.state()) <= H.REL_TOL and H.normwise(fb.cov(i), of.cov()) <= H.REL_TOL
