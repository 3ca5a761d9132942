"""GPU parity: the batched EKF-SLAM CUDA path (through the C-ABI) against the CPU oracle on identical inputs.
Tolerances are the north star's: association / landmark ids / M bit-exact; state and covariance within 1e-9
(norm-wise) per step; final trajectory within 1e-6 m / 1e-6 rad."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim():
    from live_ekf_slam_b200 import shim as s
    s.load()
    return s


def _compare(fb, inst, of, tol=H.REL_TOL):
    assert fb.num_landmarks(inst) == of.M
    assert list(fb.landmark_ids(inst)) == list(of.landmark_ids())
    ex = H.normwise(fb.state(inst), of.state())
    eP = H.normwise(fb.cov(inst), of.cov())
    assert ex <= tol and eP <= tol, (ex, eP)
    return max(ex, eP)


def test_kat_through_abi(shim, oracle):
    p = H.Params()
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 8, 4)
    fb.init(0, 0, 0)
    meas, n = fb.pack_meas([np.array([[7, 1.0, 0.0]], dtype=np.float32)])
    fb.step(0.1, 0.0, meas, n)
    d = float(np.float32(0.1))
    np.testing.assert_allclose(fb.state(0), [d, 0, 0, d + 1.0, 0], rtol=1e-15, atol=1e-18)
    P = fb.cov(0)
    np.testing.assert_allclose(P[3:, 3:], [[1.0101, 0], [0, 1.01013025]], rtol=1e-8)
    assert list(fb.landmark_ids(0)) == [7] and fb.timestep(0) == 1 and list(fb.assoc(0)) == [-1]


def test_config1_single_instance_per_step(shim, oracle):
    """BASELINE config 1: single EKF, random 20-landmark map, TSP trajectory, known IDs; checked EVERY step."""
    p, lm, fwd, ang = H.config1(seed=0, steps=1000)
    op = H.oracle_params(oracle, p)
    stream, truth = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=0, instance=0)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 20)
    of.init(0, 0, 0)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 20, 8)
    fb.init(0, 0, 0)
    worst = 0.0
    for t in range(len(fwd)):
        of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
        meas, n = fb.pack_meas([stream[t]])
        fb.step(fwd[t], ang[t], meas, n)
        assert list(fb.assoc(0)) == list(of.assoc_log()), t
        if t % 10 == 0 or t > 990:
            worst = max(worst, _compare(fb, 0, of))
    worst = max(worst, _compare(fb, 0, of))
    x, xo = fb.state(0), of.state()
    assert np.abs(x[:2] - xo[:2]).max() <= H.FINAL_TOL and abs(x[2] - xo[2]) <= H.FINAL_TOL
    assert fb.status(0) == 0 and fb.timestep(0) == len(fwd)
    assert of.M >= 15   # the TSP tour discovers (nearly) the whole map
    print("config1 worst normwise err", worst, "final M", of.M)


def test_batch_free_running_grid(shim, oracle):
    """BASELINE config 2 at reduced size: 48 instances on the 5x10 grid, 400 steps, free running."""
    p, lm, fwd, ang = H.config2(seed=1, steps=400)
    op = H.oracle_params(oracle, p)
    B = 48
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=7, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
        if t % 100 == 99:
            for i in range(0, B, 7):
                assert list(fb.assoc(i)) == list(ofs[i].assoc_log())
    worst = max(_compare(fb, i, ofs[i]) for i in range(B))
    poses = fb.poses()
    for i in range(B):
        xo = ofs[i].state()
        assert np.abs(poses[i] - xo[:3]).max() <= H.FINAL_TOL
    assert (fb.all_status() == 0).all()
    assert (fb.all_num_landmarks() == [o.M for o in ofs]).all()
    print("batch worst normwise err", worst)


def test_teacher_forced_single_steps(shim, oracle):
    """Load the oracle's (x, P, ids) of step t into the GPU filter, run ONE step, compare (SURVEY App. E protocol)."""
    p, lm, fwd, ang = H.config2(seed=2, steps=300)
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=3, instance=5)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
    of.init(0, 0, 0)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 2, 50, 8)
    fb.init(0, 0, 0)
    checked = 0
    for t in range(len(fwd)):
        if t % 17 == 0 and of.M > 0:
            fb.set_state(1, of.state(), of.cov(), of.landmark_ids(), of.timestep)
            of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
            meas, n = fb.pack_meas([[], stream[t]])
            fb.step(fwd[t], ang[t], meas, n)
            assert _compare(fb, 1, of) <= 1e-12     # one step from identical state: far tighter than 1e-9
            assert fb.timestep(1) == of.timestep
            checked += 1
        else:
            of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
    assert checked >= 10


def test_split_predict_update_matches_fused(shim, oracle):
    p, lm, fwd, ang = H.config2(seed=4, steps=120)
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=9, instance=0)
    fused = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 50, 8)
    split = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 50, 8)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
    for f in (fused, split, of):
        f.init(0, 0, 0)
    for t in range(len(fwd)):
        meas, n = fused.pack_meas([stream[t]])
        fused.step(fwd[t], ang[t], meas, n)
        split.predict(fwd[t], ang[t])
        split.update(meas, n)
        of.predict(fwd[t], ang[t])
        of.measure(stream[t])
    np.testing.assert_array_equal(fused.state(0), split.state(0))
    np.testing.assert_array_equal(fused.cov(0), split.cov(0))
    _compare(split, 0, of)
    assert split.timestep(0) == len(fwd)


def test_unknown_id_box_gate_association(shim, oracle):
    """landmark_id_is_known = false: first-match axis-aligned box gate in float (ekf.cpp:82-98); ids are slot numbers."""
    p, lm, fwd, ang = H.config2(seed=5, steps=300)
    p.landmark_id_is_known = False
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=11, instance=2)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
    of.init(0, 0, 0)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 50, 8)
    fb.init(0, 0, 0)
    matched = 0
    for t in range(len(fwd)):
        scrambled = stream[t].copy()
        if len(scrambled):
            scrambled[:, 0] = 999.0     # ids on the wire must be ignored in this mode
        of.update(fwd[t], ang[t], scrambled, oracle.DENSE)
        meas, n = fb.pack_meas([scrambled])
        fb.step(fwd[t], ang[t], meas, n)
        a = list(fb.assoc(0))
        assert a == list(of.assoc_log()), t
        matched += sum(1 for v in a if v >= 0)
    _compare(fb, 0, of)
    assert list(fb.landmark_ids(0)) == list(range(of.M))
    assert matched > 50


def test_edge_cases(shim, oracle):
    p = H.Params()
    op = H.oracle_params(oracle, p)
    # (a) no detections at all: predict-only steps (ekf.cpp:67-71)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 3, 2, 2)
    fb.init(0.5, -0.25, 0.3)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 2)
    of.init(0.5, -0.25, 0.3)
    for t in range(5):
        meas, n = fb.pack_meas([[], [], []])
        fb.step(0.05, 0.01, meas, n)
        of.update(0.05, 0.01, [])
    for i in range(3):
        _compare(fb, i, of, tol=1e-13)
    # (b) capacity: third distinct landmark is dropped and flagged
    m = np.array([[1, 1.0, 0.1], [2, 1.5, -0.2]], dtype=np.float32)
    meas, n = fb.pack_meas([m, m, m])
    fb.step(0.05, 0.0, meas, n)
    of.update(0.05, 0.0, m)
    m3 = np.array([[3, 2.0, 0.0], [1, 1.0, 0.1]], dtype=np.float32)
    meas, n = fb.pack_meas([m3, m3, m3])
    fb.step(0.05, 0.0, meas, n)
    of.update(0.05, 0.0, m3)
    assert fb.status(0) & shim.STATUS_CAPACITY and of.status & oracle.ERR_CAPACITY
    _compare(fb, 0, of)
    assert list(fb.assoc(0)) == [-1, 0]
    # (c) more detections than max_meas: flagged, extra ones dropped
    m4 = np.array([[1, 1.0, 0.1], [2, 1.5, -0.2], [1, 1.0, 0.1]], dtype=np.float32)
    meas, n = fb.pack_meas([m4, [], []])
    assert n[0] == 3
    fb.step(0.0, 0.0, meas, n)
    assert fb.status(0) & shim.STATUS_MEAS_OVERFLOW
    # (d) invalid filter choice fails loudly (localization_node.cpp:44)
    with pytest.raises(shim.SlamError):
        shim.FilterBatch(4, p.to_c(), 1, 2, 2)             # 4 = POSE_GRAPH_SLAM: not on the B200 path (localization_node.cpp:44)


def test_same_step_rematch_is_flagged(shim, oracle):
    """Unknown-ID mode: a second detection that falls inside the gate of a landmark inserted in the same step makes
    the reference index x_t out of range (ekf.cpp:115) and die; both implementations flag and freeze the instance."""
    p = H.Params()
    p.landmark_id_is_known = False
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 8, 4)
    fb.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 8)
    of.init(0, 0, 0)
    m = np.array([[0, 2.0, 0.1], [0, 2.01, 0.1]], dtype=np.float32)
    meas, n = fb.pack_meas([m])
    fb.step(0.05, 0.0, meas, n)
    of.update(0.05, 0.0, m)
    assert of.status & oracle.ERR_SAME_STEP_REMATCH
    assert fb.status(0) & shim.STATUS_SAME_STEP_REMATCH
    assert fb.num_landmarks(0) == 0 == of.M
    np.testing.assert_array_equal(fb.state(0), of.state())


def test_filter_classes_mirror_reference_interface(shim, oracle):
    """The Filter/EKF host classes: readParams -> init -> update per tick -> publishState/getStateVector."""
    from live_ekf_slam_b200.filter import make_filter, EKF, FilterChoice
    p, lm, fwd, ang = H.config1(seed=3, steps=60)
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=1, instance=0)
    filt = make_filter(p, max_landmarks=20, max_meas=8)
    assert isinstance(filt, EKF) and filt.type == FilterChoice.EKF_SLAM and not filt.isInit
    filt.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 20)
    of.init(0, 0, 0)
    for t in range(len(fwd)):
        filt.update((fwd[t], ang[t]), stream[t].reshape(-1))
        of.update(fwd[t], ang[t], stream[t])
    assert H.normwise(filt.getStateVector(), of.state()) <= H.REL_TOL
    msg = filt.publishState()
    assert msg["timestep"] == 60 and msg["M"] == of.M and msg["P"].dtype == np.float32
    assert msg["P"].size == of.n * of.n and msg["landmarks"].size == 3 * of.M
    with pytest.raises(RuntimeError):
        filt.updateNaiveVehPoseEstimate(None, None)
    bad = H.Params(filter="ekf_slam")
    bad.filter = "nope"
    with pytest.raises(RuntimeError):
        make_filter(bad)


@pytest.mark.parametrize("cap", [1, 6, 20])
def test_capacity_limited_launch_and_retry_pass(shim, oracle, cap):
    """The first pass is sized for `cap` landmarks; instances that might outgrow it are deferred untouched to the
    full-capacity retry pass.  Results must not depend on the split."""
    p, lm, fwd, ang = H.config2(seed=12, steps=220)
    op = H.oracle_params(oracle, p)
    B = 12
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.tune(0, cap)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=21, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
    for i in range(B):
        _compare(fb, i, ofs[i])
        assert fb.timestep(i) == len(fwd)
    assert max(o.M for o in ofs) > 6


@pytest.mark.parametrize("threads", [32, 64, 128, 256, 512])
def test_cta_widths_agree_with_oracle(shim, oracle, threads):
    """ekf_step_kernel is instantiated for 1..16 warps per instance (the launcher picks from the tile size); every
    width must give the oracle's results."""
    p, lm, fwd, ang = H.config2(seed=3, steps=160)
    op = H.oracle_params(oracle, p)
    B = 6
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.tune(2, threads)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=31, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
        if t % 40 == 39:
            for i in range(B):
                assert list(fb.assoc(i)) == list(ofs[i].assoc_log())
    for i in range(B):
        _compare(fb, i, ofs[i])
    assert max(o.M for o in ofs) >= 5


@pytest.mark.parametrize("threads,chunk,cap", [(0, 32, 0), (512, 1000, 0), (32, 7, 0), (64, 50, 3), (128, 16, 12), (96, 48, 0), (96, 20, 5)])
def test_sweep_kernel_matches_per_step_launches(shim, oracle, threads, chunk, cap):
    """slam_run on a known-ID EKF batch runs on the persistent ekf_sweep_kernel (simulator + filter + error terms, P
    resident in shared memory for a chunk of steps per launch, tile sized from the landmarks held so far).  Whatever
    the chunk length, the CTA width (32 .. 512 threads = 1, 2, 4, 8 filter warps; 96 = the three-warp variant that exists for
    the sweep kernel only; 0 = by tile size) and the tile capacity (cap > 0 forces small tiles: instances that outgrow them
    abort the chunk untouched and are redone by the full-capacity launch), it must reproduce the per-step launch
    sequence: same association log, landmark ids, messages, truth, statistics; state and covariance to rounding."""
    p, lm, fwd, ang = H.config2(seed=4, steps=260)
    B = 40
    res = []
    for sweep_off in (0, 1):
        fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
        fb.tune(3, sweep_off)
        if not sweep_off:
            fb.tune(2, threads)
            fb.tune(5, chunk)
            fb.tune(0, cap)
        fb.init(0, 0, 0)
        sim = shim.Simulator(fb, lm, seed=77, instance_offset=1000)
        l0 = fb.kernel_launches
        sim.run(fwd[:200], ang[:200], first_step=0)
        sim.run(fwd[200:], ang[200:], first_step=200)      # a second sweep continues from the committed state
        fb.synchronize()
        launches = fb.kernel_launches - l0
        m, n = sim.meas()
        res.append(dict(x=[fb.state(i) for i in range(B)], P=[fb.cov(i) for i in range(B)],
                        ids=[list(fb.landmark_ids(i)) for i in range(B)], assoc=[list(fb.assoc(i)) for i in range(B)],
                        ts=[fb.timestep(i) for i in range(B)], truth=sim.truth(), meas=m, n=n, stats=fb.stats(),
                        status=fb.all_status(), launches=launches))
    a, b = res
    n_chunks = -(-200 // chunk) + -(-60 // chunk)
    assert n_chunks <= a["launches"] <= 2 * n_chunks and b["launches"] >= 3 * 260
    assert a["ids"] == b["ids"] and a["assoc"] == b["assoc"] and a["ts"] == b["ts"] == [260] * B
    np.testing.assert_array_equal(a["truth"], b["truth"])
    np.testing.assert_array_equal(a["n"], b["n"])
    for i in range(B):
        np.testing.assert_array_equal(a["meas"][i, : a["n"][i]], b["meas"][i, : b["n"][i]])
        assert H.normwise(a["x"][i], b["x"][i]) <= 1e-12 and H.normwise(a["P"][i], b["P"][i]) <= 1e-12
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    np.testing.assert_allclose(a["stats"][:12], b["stats"][:12], rtol=1e-9)
    np.testing.assert_allclose(a["stats"][13], b["stats"][13], rtol=1e-9)       # same executed flops
    assert 0 < a["stats"][12] < b["stats"][12]          # the sweep kernel really moves fewer bytes: P crosses HBM once per chunk
    # and against the oracle, free running on the device simulator's own messages (1e-9 bar)
    op = H.oracle_params(oracle, p)
    msgs, _ = H.device_message_stream(shim, p, lm, fwd, ang, B, 77, 1000)
    for i in range(0, B, 9):
        filt = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        filt.init(0, 0, 0)
        for t in range(len(fwd)):
            m, n = msgs[t]
            filt.update(fwd[t], ang[t], m[i, : n[i]], oracle.STRUCTURED)
        assert filt.status == 0 and a["ids"][i] == list(filt.landmark_ids())
        assert H.normwise(a["x"][i], filt.state()) <= H.REL_TOL and H.normwise(a["P"][i], filt.cov()) <= H.REL_TOL


def test_sweep_kernel_freezes_dead_instance(shim, oracle):
    """A repeated id inside one message makes the reference die (ekf.cpp:115).  The per-step path flags it and freezes
    the instance at its last committed state; a sweep must do the same (here: duplicate landmark positions in the map
    never produce duplicate ids, so the dead path is driven through slam_step and the sweep continues frozen)."""
    p, lm, fwd, ang = H.config2(seed=5, steps=60)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 2, 50, 8)
    fb.init(0, 0, 0)
    meas, n = fb.pack_meas([np.array([[3, 1.0, 0.1], [3, 1.0, 0.1]], dtype=np.float32), np.zeros((0, 3), dtype=np.float32)])
    fb.step(0.05, 0.0, meas, n)
    assert fb.status(0) & shim.STATUS_SAME_STEP_REMATCH and fb.status(1) == 0
    x0, P0 = fb.state(0), fb.cov(0)
    sim = shim.Simulator(fb, lm, seed=1)
    sim.run(fwd, ang)
    np.testing.assert_array_equal(fb.state(0), x0)
    np.testing.assert_array_equal(fb.cov(0), P0)
    assert fb.timestep(0) == 0 and fb.timestep(1) == 61 and fb.num_landmarks(0) == 0


@pytest.mark.parametrize("chunk,cap,poses", [(32, 0, True), (9, 4, True), (1000, 0, False)])
def test_run_io_replays_recorded_messages(shim, oracle, chunk, cap, poses):
    """slam_run_io: a whole recorded run through HOST buffers (chunks uploaded / filtered by the replay mode of
    ekf_sweep_kernel / poses downloaded, pipelined).  Must equal T calls of slam_step + the pose read-back."""
    p, lm, fwd, ang = H.config2(seed=6, steps=150)
    op = H.oracle_params(oracle, p)
    B, T, mm = 10, len(fwd), 8
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=5, instance=i)[0] for i in range(B)]
    meas = np.zeros((T, B, mm, 3), dtype=np.float32)
    nm = np.zeros((T, B), dtype=np.int32)
    ref = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, mm)
    ref.init(0, 0, 0)
    ref_poses = np.zeros((T, B, 3))
    for t in range(T):
        meas[t], nm[t] = ref.pack_meas([streams[i][t] for i in range(B)])
        ref.step(fwd[t], ang[t], meas[t], nm[t])
        ref_poses[t] = ref.poses()
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, mm)
    fb.tune(5, chunk)
    fb.tune(0, cap)
    fb.init(0, 0, 0)
    out = np.full((T, B, 3), np.nan) if poses else None
    half = 70
    fb.run_io(fwd[:half], ang[:half], 0, meas[:half], nm[:half], out[:half] if poses else None, half)
    fb.run_io(fwd[half:], ang[half:], 0, meas[half:], nm[half:], out[half:] if poses else None, T - half)
    fb.synchronize()
    if poses:
        assert np.abs(out - ref_poses).max() <= 1e-12
    for i in range(B):
        assert list(fb.landmark_ids(i)) == list(ref.landmark_ids(i)) and list(fb.assoc(i)) == list(ref.assoc(i))
        assert fb.timestep(i) == T
        assert H.normwise(fb.state(i), ref.state(i)) <= 1e-12 and H.normwise(fb.cov(i), ref.cov(i)) <= 1e-12
    of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
    of.init(0, 0, 0)
    for t in range(T):
        of.update(fwd[t], ang[t], streams[3][t], oracle.STRUCTURED)
    _compare(fb, 3, of)


def test_run_io_generic_path_unknown_ids(shim, oracle):
    """slam_run_io for a filter the sweep kernel does not cover (unknown-ID association): chunked uploads feeding
    per-step launches."""
    p, lm, fwd, ang = H.config2(seed=8, steps=90)
    p.landmark_id_is_known = False
    op = H.oracle_params(oracle, p)
    B, T, mm = 3, len(fwd), 8
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=2, instance=i)[0] for i in range(B)]
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, mm)
    fb.tune(5, 25)
    fb.init(0, 0, 0)
    meas = np.zeros((T, B, mm, 3), dtype=np.float32)
    nm = np.zeros((T, B), dtype=np.int32)
    for t in range(T):
        meas[t], nm[t] = fb.pack_meas([streams[i][t] for i in range(B)])
    out = np.zeros((T, B, 3))
    fb.run_io(fwd, ang, 0, meas, nm, out, T)
    fb.synchronize()
    for i in range(B):
        of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        of.init(0, 0, 0)
        for t in range(T):
            of.update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
        _compare(fb, i, of)
        assert np.abs(out[-1, i] - of.state()[:3]).max() <= H.FINAL_TOL


@pytest.mark.parametrize("kind,B", [("ekf", 12), ("ukf", 12), ("ekf", 80)], ids=["ekf", "ukf", "ekf_gathered_inputs"])
def test_step_io_zero_copy_matches_staged(shim, oracle, kind, B):
    """slam_step_io with PINNED host buffers: the kernels read the command / message buffers and write the poses in place
    over PCIe (batched EKF: pose write fused into the step kernel, one launch per tick).  Must equal the staged path
    (slam_tune key 14) and pageable numpy buffers, tick by tick."""
    import torch
    filt = "ekf_slam" if kind == "ekf" else "ukf_slam"
    p, lm, fwd, ang = H.config2(seed=9, steps=70, filt=filt)
    op = H.oracle_params(oracle, p)
    T, mm = len(fwd), 8
    skind = shim.EKF_SLAM if kind == "ekf" else shim.UKF_SLAM
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=4, instance=i)[0] for i in range(B)]
    probe = shim.FilterBatch(skind, p.to_c(), B, 50, mm)
    h_meas = torch.zeros((T, B, mm, 3), dtype=torch.float32).pin_memory()
    h_n = torch.zeros((T, B), dtype=torch.int32).pin_memory()
    for t in range(T):
        m, n = probe.pack_meas([streams[i][t] for i in range(B)])
        h_meas[t].copy_(torch.from_numpy(m)); h_n[t].copy_(torch.from_numpy(n))
    h_fwd = torch.from_numpy(fwd.copy()).pin_memory()
    h_ang = torch.from_numpy(ang.copy()).pin_memory()
    runs = []
    for mode in ("zero_copy", "staged", "pageable"):
        fb = shim.FilterBatch(skind, p.to_c(), B, 50, mm)
        fb.tune(14, 1 if mode == "staged" else 0)
        fb.init(0, 0, 0)
        if mode == "pageable":
            poses = np.zeros((T, B, 3))
            l0 = fb.kernel_launches
            for t in range(T):
                fb.step_io(fwd[t:t + 1].copy(), ang[t:t + 1].copy(), 0, h_meas[t].numpy().copy(), h_n[t].numpy().copy(), poses[t])
                fb.synchronize()
        else:
            h_pose = torch.full((T, B, 3), float("nan"), dtype=torch.float64).pin_memory()
            l0 = fb.kernel_launches
            for t in range(T):
                fb.step_io(h_fwd.data_ptr() + 4 * t, h_ang.data_ptr() + 4 * t, 0, h_meas[t].data_ptr(), h_n[t].data_ptr(), h_pose[t].data_ptr())
                fb.synchronize()
                assert torch.isfinite(h_pose[t]).all()        # readable right after the tick's synchronize
            poses = h_pose.numpy().copy()
        runs.append(dict(poses=poses, x=[fb.state(i) for i in range(B)], P=[fb.cov(i) for i in range(B)], launches=fb.kernel_launches - l0))
    for other in runs[1:]:
        np.testing.assert_array_equal(runs[0]["poses"], other["poses"])
        for i in range(B):
            np.testing.assert_array_equal(runs[0]["x"][i], other["x"][i])
            np.testing.assert_array_equal(runs[0]["P"][i], other["P"][i])
    if kind == "ekf" and B < 64:
        assert runs[0]["launches"] < runs[1]["launches"]       # no separate pose kernel on the zero-copy path (small batch: inputs read in place)
    okind = oracle.EKF_SLAM if kind == "ekf" else oracle.UKF_SLAM
    of = oracle.OracleFilter(okind, op, 50)
    of.init(0, 0, 0)
    for t in range(T):
        of.update(fwd[t], ang[t], streams[5][t], oracle.STRUCTURED)
        xo = of.state()
        yaw = xo[2] if kind == "ekf" else np.arctan2(xo[3], xo[2])
        assert np.abs(runs[0]["poses"][t, 5] - [xo[0], xo[1], yaw]).max() <= H.FINAL_TOL
