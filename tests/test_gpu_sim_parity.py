"""GPU parity of the workload source: the on-GPU measurement generator (sim_node.py:209-250 restated in
csrc/sim.cu) against the CPU oracle, and the fused sim+filter sweep (slam_run) against oracle_run_instance."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim():
    from live_ekf_slam_b200 import shim as s
    s.load()
    return s


def test_sim_messages_bit_exact(shim, oracle):
    p, lm, fwd, ang = H.config2(seed=0, steps=250)
    op = H.oracle_params(oracle, p)
    B, seed, off = 16, 1234567890123, 3
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    sim = shim.Simulator(fb, lm, seed=seed, instance_offset=off)
    truths = [np.zeros(3) for _ in range(B)]
    n_msgs = 0
    ulp_flips = 0
    for t in range(len(fwd)):
        sim.step(fwd[t], ang[t], t)
        m, n = sim.meas()
        tr = sim.truth()
        for i in range(B):
            ref = oracle.sim_step(op, truths[i], fwd[t], ang[t], lm, seed, off + i, t)
            assert n[i] == len(ref), (t, i)
            got = m[i, : n[i]]
            np.testing.assert_array_equal(got[:, 0], ref[:, 0])          # ids and visibility decisions: exact
            if not np.array_equal(got, ref):
                # device libm differs from glibc by <= 2 ulp(double); after rounding to float32 that flips at most
                # the last float32 bit, on ~1e-8 of the values.
                np.testing.assert_allclose(got, ref, rtol=1.3e-7, atol=0)
                ulp_flips += int((got != ref).sum())
            n_msgs += len(ref)
            assert np.abs(tr[i] - truths[i]).max() <= 1e-12
    assert n_msgs > 3000
    assert ulp_flips <= 2, ulp_flips


def test_slam_run_matches_oracle_instances(shim, oracle):
    """The device-side sweep (sim -> filter -> error accumulators) reproduces independent oracle instances."""
    p, lm, fwd, ang = H.config2(seed=6, steps=300)
    op = H.oracle_params(oracle, p)
    B, seed, off = 24, 42, 100
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.init(0, 0, 0)
    sim = shim.Simulator(fb, lm, seed=seed, instance_offset=off)
    sim.run(fwd, ang, first_step=0)
    poses = fb.poses()
    truth = sim.truth()
    sum_pos_err = 0.0
    sum_sq = np.zeros(3)
    # the oracle filter is fed the device simulator's own messages (twin handle), so state and covariance are held to
    # the 1e-9 bar; the fully independent oracle run (its own simulator) pins the trajectory and the truth
    msgs, dev_truth = H.device_message_stream(shim, p, lm, fwd, ang, B, seed, off)
    for i in range(B):
        st, pose_ind, tr, _ = oracle.run_instance(oracle.EKF_SLAM, op, lm, fwd, ang, seed, off + i, 50, oracle.STRUCTURED)
        assert st == 0
        assert np.abs(poses[i] - pose_ind[-1]).max() <= H.FINAL_TOL
        assert np.abs(truth[i] - tr[-1]).max() <= 1e-11 and np.abs(dev_truth[:, i] - tr).max() <= 1e-11
        filt = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        filt.init(0, 0, 0)
        pose = np.zeros((len(fwd), 3))
        for t in range(len(fwd)):
            m, n = msgs[t]
            filt.update(fwd[t], ang[t], m[i, : n[i]], oracle.STRUCTURED)
            pose[t] = filt.state()[:3]
        assert fb.num_landmarks(i) == filt.M
        assert list(fb.landmark_ids(i)) == list(filt.landmark_ids())
        assert H.normwise(fb.state(i), filt.state()) <= H.REL_TOL and H.normwise(fb.cov(i), filt.cov()) <= H.REL_TOL
        e = pose - tr
        e[:, 2] = np.remainder(e[:, 2] + np.pi, 2 * np.pi) - np.pi
        sum_pos_err += np.sqrt(e[:, 0] ** 2 + e[:, 1] ** 2).sum()
        sum_sq += (e ** 2).sum(axis=0)
    s = fb.stats()
    assert s[0] == B * len(fwd)
    np.testing.assert_allclose(s[1:4], sum_sq, rtol=1e-6)
    np.testing.assert_allclose(s[4], sum_pos_err, rtol=1e-6)
    assert s[5] > 0 and s[6] == 0 and s[7] == fb.all_num_landmarks().sum()


def test_shard_invariance(shim, oracle):
    """Instance i gives bit-identical results whether it runs in a batch of 16 at offset 0 or in a shard of 4 at
    offset 8 (multi-GPU sharding changes nothing: RNG is keyed by the global instance id)."""
    p, lm, fwd, ang = H.config2(seed=8, steps=150)
    full = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 16, 50, 8)
    sim_full = shim.Simulator(full, lm, seed=5, instance_offset=0)
    sim_full.run(fwd, ang)
    shard = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 4, 50, 8)
    sim_shard = shim.Simulator(shard, lm, seed=5, instance_offset=8)
    sim_shard.run(fwd, ang)
    for j in range(4):
        np.testing.assert_array_equal(full.state(8 + j), shard.state(j))
        np.testing.assert_array_equal(full.cov(8 + j), shard.cov(j))


def test_sim_wide_kernel_large_map(shim, oracle):
    """Few vehicles on a large map take the CTA-per-vehicle generator (32 warps, chunk counts scanned in shared memory):
    same messages, in ascending-id order, as the oracle; overflow beyond max_meas truncates identically."""
    from live_ekf_slam_b200 import workload as wl
    p = H.Params(filter="ekf_slam")
    op = H.oracle_params(oracle, p)
    rng = np.random.default_rng(5)
    lm = wl.random_map_fast(1500, p.map_bound, 0.3, rng)
    fwd, ang = wl.tsp_trajectory(lm, p, rng, 60)
    B, seed = 2, 99
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 64)
    sim = shim.Simulator(fb, lm, seed=seed, instance_offset=0)
    truths = [np.zeros(3) for _ in range(B)]
    n_msgs, flips = 0, 0
    for t in range(len(fwd)):
        sim.step(fwd[t], ang[t], t)
        m, n = sim.meas()
        tr = sim.truth()
        for i in range(B):
            ref = oracle.sim_step(op, truths[i], fwd[t], ang[t], lm, seed, i, t)[:64]
            assert n[i] == len(ref), (t, i, n[i], len(ref))
            got = m[i, : n[i]]
            np.testing.assert_array_equal(got[:, 0], ref[:, 0])
            if not np.array_equal(got, ref):
                np.testing.assert_allclose(got, ref, rtol=1.3e-7, atol=0)
                flips += int((got != ref).sum())
            n_msgs += len(ref)
            assert np.abs(tr[i] - truths[i]).max() <= 1e-12
    assert n_msgs > 2000 and flips <= 2


def test_error_histogram_vs_oracle_runs(shim, oracle):
    """slam_get_error_histogram: the per-run average position error (plotting_node.py:195-218) of every instance against
    oracle runs of the same instances, and the on-device histogram against numpy on those numbers (exact counts)."""
    p, lm, fwd, ang = H.config2(seed=1, steps=150)
    op = H.oracle_params(oracle, p)
    B, seed = 40, 77
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.init(0, 0, 0)
    sim = shim.Simulator(fb, lm, seed=seed)
    sim.run(fwd, ang)
    fb.synchronize()
    lo, hi, nb = 0.0, 0.5, 25
    counts, avg = fb.error_histogram(lo, hi, nb, per_instance=True)
    for i in (0, 7, 39):
        st, pose, truth, _ = oracle.run_instance(oracle.EKF_SLAM, op, lm, fwd, ang, seed, i, 50, oracle.STRUCTURED)
        ref = np.sqrt(((pose[:, :2] - truth[:, :2]) ** 2).sum(axis=1)).mean()
        assert abs(avg[i] - ref) <= 1e-9 * max(1.0, ref), (i, avg[i], ref)
    ref_counts = np.zeros(nb + 2, dtype=np.int64)
    for v in avg:
        k = 0 if v < lo else (nb + 1 if v >= hi else 1 + min(int((v - lo) / (hi - lo) * nb), nb - 1))
        ref_counts[k] += 1
    np.testing.assert_array_equal(counts, ref_counts)
    assert counts.sum() == B and abs(avg.mean() - fb.stats()[4] / fb.stats()[0]) <= 1e-12
    with pytest.raises(shim.SlamError):
        fb.error_histogram(1.0, 1.0, 4)


def test_device_tsp_trajectories(shim, oracle):
    """slam_sim_make_trajectories: per-instance command trajectories generated on the device (sim_node.py:63-152) against
    the CPU restatement, then a sweep driven by them (cmd_stride = 1) against oracle runs with the same commands."""
    import torch
    p, lm, _, _ = H.config2(seed=0, steps=10)
    op = H.oracle_params(oracle, p)
    B, T, seed, off = 24, 300, 4242, 5
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.init(*p.init_pose)
    sim = shim.Simulator(fb, lm, seed=seed, instance_offset=off)
    d_fwd = torch.zeros((T, B), dtype=torch.float32, device="cuda")
    d_ang = torch.zeros((T, B), dtype=torch.float32, device="cuda")
    sim.make_trajectories(p.landmark_noise, p.visitation_threshold, p.map_bound, p.init_pose, T, d_fwd, d_ang)
    fb.synchronize()
    fwd, ang = d_fwd.cpu().numpy(), d_ang.cpu().numpy()
    flips = 0
    refs = {}
    for i in range(B):
        rf, ra = oracle.tsp_trajectory(op, lm, p.landmark_noise, p.visitation_threshold, p.map_bound, p.init_pose, T, seed, off + i)
        refs[i] = (rf, ra)
        assert np.abs(fwd[:, i] - rf).max() <= 1e-7 and np.abs(ang[:, i] - ra).max() <= 1e-7, i
        # while the vehicle converges on its goal the heading command is the residual of a cancellation (gb - th): its
        # absolute error stays ~1e-16 (device vs glibc atan2, FMA contraction) whatever its size; commands of visible size
        # may only flip a last float32 bit
        np.testing.assert_allclose(ang[:, i], ra, rtol=2.5e-7, atol=1e-12)
        np.testing.assert_allclose(fwd[:, i], rf, rtol=2.5e-7, atol=0)
        big = np.abs(ra) > 1e-6
        flips += int((fwd[:, i] != rf).sum() + (ang[big, i] != ra[big]).sum())
    assert flips <= 6, flips
    assert not np.array_equal(ang[:, 0], ang[:, 1])
    sim.run_device(d_fwd, d_ang, 1, T, 0)
    fb.synchronize()
    for i in (0, 11, 23):
        st, pose, truth, keep = oracle.run_instance(oracle.EKF_SLAM, op, lm, fwd[:, i].copy(), ang[:, i].copy(), seed, off + i, 50,
                                                    oracle.STRUCTURED, keep=True)
        assert fb.num_landmarks(i) == keep.M and list(fb.landmark_ids(i)) == list(keep.landmark_ids())
        assert H.normwise(fb.state(i), keep.state()) <= H.REL_TOL and H.normwise(fb.cov(i), keep.cov()) <= H.REL_TOL
