"""GPU parity: the batched UKF-SLAM CUDA path (through the C-ABI) against the CPU oracle on identical inputs.
Tolerances: landmark ids / M / association bit-exact; state and covariance within 1e-9 norm-wise per step
(|delta| <= 1e-9 * max(1, max|.|), SURVEY.md 7 hard part 2); final pose within 1e-6 m / 1e-6 rad."""
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


def test_ukf_kats_through_abi(shim, oracle):
    p = H.Params(filter="ukf_slam")
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 1, 8, 4)
    fb.init(1.0, 2.0, 0.0)
    meas, n = fb.pack_meas([[]])
    fb.step(0.0, 0.0, meas, n)
    x = fb.state(0)
    sw = float(np.float32(0.2)) + 8 * float(np.float32(0.1))       # KAT-3: weights do not sum to 1
    assert abs(x[0] - sw * 1.0) < 1e-12 and abs(x[1] - sw * 2.0) < 1e-12
    P = fb.cov(0)
    assert abs(P[0, 0] - (1e-4 + 0.01)) < 1e-9 and abs(P[1, 1] - 1e-4) < 1e-9   # first-step Q = diag(.01,0,.01,0)
    assert fb.timestep(0) == 1


def test_ukf_single_instance_per_step(shim, oracle):
    """UKF-SLAM on the 5x10 grid, every step compared with the oracle (free running, 500 steps)."""
    p, lm, fwd, ang = H.config2(seed=0, steps=500, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=0, instance=0)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    of.init(0, 0, 0)
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 1, 50, 8)
    fb.init(0, 0, 0)
    worst = 0.0
    for t in range(len(fwd)):
        of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
        meas, n = fb.pack_meas([stream[t]])
        fb.step(fwd[t], ang[t], meas, n)
        assert list(fb.assoc(0)) == list(of.assoc_log()), t
        if t % 5 == 0 or t > 490:
            worst = max(worst, _compare(fb, 0, of))
    xv = fb.state_vector(0)
    xo = of.state()
    yaw_o = np.arctan2(xo[3], xo[2])
    assert np.abs(xv[:2] - xo[:2]).max() <= H.FINAL_TOL and abs(xv[2] - yaw_o) <= H.FINAL_TOL
    assert fb.status(0) == 0 and fb.timestep(0) == len(fwd) and of.M >= 20
    print("ukf single worst normwise err", worst, "final M", of.M)


def test_ukf_batch_free_running(shim, oracle):
    p, lm, fwd, ang = H.config2(seed=3, steps=250, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    B = 12
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), B, 50, 8)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=5, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
    worst = max(_compare(fb, i, ofs[i]) for i in range(B))
    poses = fb.poses()
    for i in range(B):
        xo = ofs[i].state()
        assert np.abs(poses[i, :2] - xo[:2]).max() <= H.FINAL_TOL
        assert abs(poses[i, 2] - np.arctan2(xo[3], xo[2])) <= H.FINAL_TOL
    assert (fb.all_status() == 0).all()
    print("ukf batch worst normwise err", worst)


def test_ukf_teacher_forced_single_steps(shim, oracle):
    p, lm, fwd, ang = H.config2(seed=2, steps=260, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=3, instance=5)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    of.init(0, 0, 0)
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 2, 50, 8)
    fb.init(0, 0, 0)
    checked, worst = 0, 0.0
    for t in range(len(fwd)):
        if t % 13 == 0 and of.M > 0:
            fb.set_state(1, of.state(), of.cov(), of.landmark_ids(), of.timestep)
            of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
            meas, n = fb.pack_meas([[], stream[t]])
            fb.step(fwd[t], ang[t], meas, n)
            worst = max(worst, _compare(fb, 1, of))
            checked += 1
        else:
            of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
    assert checked >= 10
    print("ukf teacher-forced worst", worst)


@pytest.mark.parametrize("knobs", [(), ((17, 0),), ((17, 1),), ((15, 0),), ((15, 2),), ((10, 1),), ((7, 2),)],
                         ids=["default", "full_square_tridiagonalisation", "packed_tridiagonalisation", "global_scratch_eigenvectors",
                              "tile_then_global_scratch", "one_slice", "generation2"])
def test_ukf_teacher_forced_large_states(shim, oracle, knobs):
    """Single steps from the oracle's state LATE in a run (40+ landmarks, n = 84 .. 104): the sizes where the step takes the packed
    tridiagonalisation, the two- and three-per-SM classes of the eigenvector tile and the largest reflector / dense-product loops
    -- none of which the short free-running tests reach.  The oracle walks the run in STRUCTURED mode (cross-checked against the
    dense-faithful mode elsewhere) and the compared step itself is dense-faithful."""
    p, lm, fwd, ang = H.config2(seed=7, steps=1300, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=11, instance=2)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    of.init(0, 0, 0)
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 3, 50, 8)
    for k, v in knobs:
        fb.tune(k, v)
    fb.init(0, 0, 0)
    checked, worst, sizes = 0, 0.0, set()
    for t in range(len(fwd)):
        if of.M >= 40 and t % 61 == 0 and checked < 8:
            for i in range(3):
                fb.set_state(i, of.state(), of.cov(), of.landmark_ids(), of.timestep)
            of.update(fwd[t], ang[t], stream[t], oracle.DENSE)
            meas, n = fb.pack_meas([stream[t], [], stream[t]])
            fb.step(fwd[t], ang[t], meas, n)
            worst = max(worst, _compare(fb, 0, of), _compare(fb, 2, of))
            sizes.add(4 + 2 * of.M)
            checked += 1
        else:
            of.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
    assert checked >= 6 and min(sizes) <= 90 and max(sizes) == 104, (checked, sizes)
    assert (fb.all_status() == 0).all()
    print("ukf teacher-forced large states", knobs, sorted(sizes), "worst", worst)


def test_ukf_edge_cases(shim, oracle):
    p = H.Params(filter="ukf_slam")
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 2, 2, 3)
    fb.init(0.5, -0.25, 0.3)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 2)
    of.init(0.5, -0.25, 0.3)
    for _ in range(4):                                  # predict-only steps
        meas, n = fb.pack_meas([[], []])
        fb.step(0.05, 0.01, meas, n)
        of.update(0.05, 0.01, [])
    _compare(fb, 0, of)
    m = np.array([[1, 1.0, 0.1], [2, 1.5, -0.2], [3, 2.0, 0.0]], dtype=np.float32)   # third insertion exceeds capacity
    meas, n = fb.pack_meas([m, m])
    fb.step(0.05, 0.0, meas, n)
    of.update(0.05, 0.0, m)
    assert fb.status(0) & shim.STATUS_CAPACITY and of.status & oracle.ERR_CAPACITY
    _compare(fb, 1, of)
    m2 = np.array([[2, 1.4, -0.25], [1, 0.9, 0.12]], dtype=np.float32)               # two updates, message order
    meas, n = fb.pack_meas([m2, m2])
    fb.step(0.05, 0.0, meas, n)
    of.update(0.05, 0.0, m2)
    _compare(fb, 0, of)
    assert list(fb.assoc(0)) == [1, 0]
    with pytest.raises(shim.SlamError):                 # split predict/update is EKF-only (ukf.cpp:305-337)
        fb.predict(0.1, 0.0)


def test_ukf_filter_class(shim, oracle):
    from live_ekf_slam_b200.filter import make_filter, UKF
    p, lm, fwd, ang = H.config2(seed=3, steps=40, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=1, instance=0)
    filt = make_filter(p, max_landmarks=50, max_meas=8)
    assert isinstance(filt, UKF)
    filt.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    of.init(0, 0, 0)
    for t in range(len(fwd)):
        filt.update((fwd[t], ang[t]), stream[t].reshape(-1))
        of.update(fwd[t], ang[t], stream[t])
    xo = of.state()
    sv = filt.getStateVector()                            # (x, y, yaw, landmarks...) -- the vector ukf.cpp:47-53 means to build
    assert sv.size == 3 + 2 * of.M
    assert H.normwise(sv[3:], xo[4:]) <= H.REL_TOL and abs(sv[2] - np.arctan2(xo[3], xo[2])) <= 1e-9
    msg = filt.publishState()
    assert msg["P"].size == of.n * of.n and msg["M"] == of.M


@pytest.mark.parametrize("knobs", [((7, 1),), ((7, 2),), ((7, 2), (8, 600)), ((7, 2), (9, 1)), ((7, 2), (8, 2500), (9, 1)), ((7, 2), (11, 0)),
                                   ((7, 3),), ((7, 3), (12, 1)), ((7, 3), (12, 1), (8, 600)), ((7, 3), (9, 1)), ((7, 3), (11, 0)),
                                   ((7, 3), (13, 0)), ((7, 3), (13, 0), (9, 1)), ((7, 3), (15, 0)), ((7, 3), (15, 0), (12, 1)), ((7, 3), (16, 1)), ((7, 3), (11, 2)), ((7, 3), (17, 0)), ((7, 2), (17, 0)), ((7, 3), (17, 1))],
                         ids=["generation1", "generation2", "rescue_pass_only", "clip_overflow_pass", "mixed_rescue", "full_width_tile_only",
                              "generation3", "gen3_clusters_take_the_ql_route", "gen3_ql_route_then_rescue", "gen3_clip_overflow_pass",
                              "gen3_full_width_tile_only", "gen3_single_warp_back_kernel", "gen3_single_warp_clip_overflow",
                              "gen3_global_scratch_eigenvectors", "gen3_global_scratch_ql_route", "gen3_tile_hands_over_clusters", "gen3_three_tile_widths",
                              "gen3_full_square_tridiagonalisation", "gen2_full_square_tridiagonalisation", "gen3_packed_tridiagonalisation"])
def test_ukf_step_variants(shim, oracle, knobs):
    """The same free-running batch through the alternative code paths of the UKF step: the generation-1 kernels
    (explicit eigenvectors), a rotation log too small for any / for the later steps (rescue pass on the generation-1
    kernels), only one clipped eigenvector allowed beside the first S-pass (overflow pass into the seed); generation 3
    (parallel tridiagonal eigensolver + dense products, the default) alone, with every cluster of close eigenvalues declined
    (those instances take the QL route of generation 2 in the same step), and with that route's log too small as well."""
    p, lm, fwd, ang = H.config2(seed=4, steps=120, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    B = 5
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), B, 50, 8)
    for k, v in knobs:
        fb.tune(k, v)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=9, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
            if t % 10 == 0:
                assert list(fb.assoc(i)) == list(ofs[i].assoc_log()), (t, i)
    worst = max(_compare(fb, i, ofs[i]) for i in range(B))
    assert (fb.all_status() == 0).all() and ofs[0].M >= 5
    print("ukf variant", knobs, "worst normwise err", worst)


def test_ukf_sliced_batch(shim, oracle):
    """The batch cut into slices that run front -> QL -> back on separate streams (slam_tune key 10): instances on both
    sides of a slice boundary against the oracle, and the whole batch against the unsliced run."""
    p, lm, fwd, ang = H.config2(seed=6, steps=70, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    B = 150
    runs = []
    for nsub in (1, 2):
        fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), B, 50, 8)
        fb.tune(10, nsub)
        fb.init(0, 0, 0)
        sim = shim.Simulator(fb, lm, seed=21)
        sim.run(fwd, ang)
        fb.synchronize()
        runs.append(fb)
    for i in (0, 74, 75, 149):
        stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=21, instance=i)
        of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
        of.init(0, 0, 0)
        for t in range(len(fwd)):
            of.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
        _compare(runs[1], i, of)
    for i in range(0, B, 7):        # (the sliced path runs the shared-memory-free QL instantiation: same recurrence, own FMA contraction)
        assert H.normwise(runs[0].state(i), runs[1].state(i)) <= 1e-10 and H.normwise(runs[0].cov(i), runs[1].cov(i)) <= 1e-10
    assert (runs[1].all_status() == 0).all()


def test_ukf_narrow_tile_hand_over(shim, oracle):
    """Steps with more than four updates do not fit the narrow tile of the back kernel: the instance flags itself and the
    full-width pass of the same step takes it; its neighbour (fewer updates) stays on the narrow pass."""
    p = H.Params(filter="ukf_slam")
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 2, 10, 8)
    fb.init(0, 0, 0)
    ofs = [oracle.OracleFilter(oracle.UKF_SLAM, op, 10) for _ in range(2)]
    for of in ofs:
        of.init(0, 0, 0)
    six = np.array([[i, 1.0 + 0.2 * i, -0.5 + 0.2 * i] for i in range(6)], dtype=np.float32)
    script = [(six, six[:2]),                       # step 1: insertions only
              (six + np.float32([0, 0.01, -0.01]), six[:2]),      # step 2: six updates (hand-over) / two updates (narrow)
              (six[1:4], six[:1]),                  # step 3: three updates / one
              (six[::-1].copy(), np.zeros((0, 3), np.float32))]   # step 4: six updates in reverse message order / none
    for ma, mb in script:
        meas, n = fb.pack_meas([ma, mb])
        fb.step(0.05, 0.01, meas, n)
        ofs[0].update(0.05, 0.01, ma)
        ofs[1].update(0.05, 0.01, mb)
        for i in range(2):
            assert list(fb.assoc(i)) == list(ofs[i].assoc_log())
            _compare(fb, i, ofs[i])
    assert (fb.all_status() == 0).all() and fb.num_landmarks(0) == 6


@pytest.mark.parametrize("max_lm,max_meas", [(70, 8), (70, 16)], ids=["n_max_144_generation2", "max_meas_16_generation1_fallback"])
def test_ukf_capacity_variants(shim, oracle, max_lm, max_meas):
    """Handle capacities beyond the benchmark's: n_max = 144 takes the 8-row-slot reflector products of generation 2;
    max_meas = 16 exceeds the lanes of the generation-2 tile, so the library falls back to the generation-1 kernels."""
    p, lm, fwd, ang = H.config2(seed=8, steps=90, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    B = 3
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), B, max_lm, max_meas)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=12, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.UKF_SLAM, op, max_lm)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
    worst = max(_compare(fb, i, ofs[i]) for i in range(B))
    assert (fb.all_status() == 0).all() and ofs[0].M >= 4
    print("ukf capacity variant", max_lm, max_meas, "worst", worst)


@pytest.mark.parametrize("knobs", [(), ((7, 2),), ((7, 1),), ((7, 2), (8, 600)), ((7, 3), (12, 1))],
                         ids=["generation3", "generation2", "generation1", "rescue_pass", "gen3_ql_route"])
def test_ukf_sigma_points_getter(shim, oracle, knobs):
    """slam_get_sigma_points: X of the last update (ukf.cpp:214-220, published point-major by ukf.cpp:91-99), materialised
    on demand from the factors the step kernels leave on the device, against the oracle's X after the same update.
    X keeps the size of the step's PRIOR (ukf.cpp:167-171) even when the step inserted landmarks."""
    p, lm, fwd, ang = H.config2(seed=5, steps=140, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    B = 3
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), B, 50, 8)
    for k, v in knobs:
        fb.tune(k, v)
    fb.init(0, 0, 0)
    X0 = fb.sigma_points(1)
    assert X0.shape == (9, 4) and not X0.any()                  # the constructor's zero matrix (ukf.cpp:20)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=13, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    checked, grew, worst = 0, 0, 0.0
    for t in range(len(fwd)):
        n_before = [o.n for o in ofs]
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.STRUCTURED)
        if t % 9 == 0 or t == len(fwd) - 1 or any(o.n != nb for o, nb in zip(ofs, n_before)):
            for i in range(B):
                X, Xo = fb.sigma_points(i), ofs[i].sigma_points()
                assert X.shape == Xo.shape == (2 * n_before[i] + 1, n_before[i]), (t, i, X.shape, Xo.shape)
                e = H.normwise(X, Xo)
                assert e <= H.REL_TOL, (t, i, e)
                worst = max(worst, e)
                checked += 1
                grew += int(ofs[i].n != n_before[i])
    assert checked > 40 and grew > 5 and ofs[0].M >= 5
    # the wire layout: UKFState.X is X flattened column by column (ukf.cpp:93-97)
    from live_ekf_slam_b200.filter import make_filter
    filt = make_filter(p, max_landmarks=50, max_meas=8)
    filt.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    of.init(0, 0, 0)
    for t in range(40):
        filt.update((fwd[t], ang[t]), streams[0][t].reshape(-1))
        of.update(fwd[t], ang[t], streams[0][t], oracle.STRUCTURED)
    msg = filt.publishState()
    Xo = of.sigma_points()
    assert msg["X"].dtype == np.float32 and msg["X"].size == Xo.size
    np.testing.assert_allclose(msg["X"], Xo.reshape(-1).astype(np.float32), rtol=2e-7, atol=1e-9)
    assert filt.X.shape == (Xo.shape[1], Xo.shape[0])
    print("ukf sigma points worst normwise err", worst, "checked", checked)


def test_ukf_loc_sigma_points(shim, oracle):
    p, lm, fwd, ang = H.config2(seed=2, steps=40, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(shim.UKF_LOC, p.to_c(), 2, len(lm), 8)
    fb.set_map(lm)
    fb.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_LOC, op, len(lm))
    of.init(0, 0, 0)
    of.set_map(lm)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=3, instance=0)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([stream[t], stream[t]])
        fb.step(fwd[t], ang[t], meas, n)
        of.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
        X, Xo = fb.sigma_points(1), of.sigma_points()
        assert X.shape == Xo.shape == (9, 4)
        assert H.normwise(X, Xo) <= H.REL_TOL


def test_ukf_many_clipped_eigenvalues_take_the_rescue_pass(shim, oracle):
    """nearestSPD clips ANY number of non-positive eigenvalues (ukf.cpp:120).  The warp-per-instance kernel carries at most 32
    clipped eigenvectors; an instance with more is left untouched and redone by the generation-1 rescue pass in the same step
    (it used to be flagged NaN).  Teacher-forced: a covariance whose 40 landmark directions are all negative."""
    p = H.Params(filter="ukf_slam")
    op = H.oracle_params(oracle, p)
    M = 20
    n = 4 + 2 * M
    rng = np.random.default_rng(2)
    x = np.concatenate([[0.3, -0.2, np.cos(0.4), np.sin(0.4)], rng.uniform(-3, 3, 2 * M)])
    A = rng.normal(size=(n, n)) * 1e-3
    P = -(A @ A.T) - 1e-4 * np.eye(n)                     # negative definite: every eigenvalue is clipped to 1e-8
    P[:4, :4] = np.diag([1e-2, 1e-2, 1e-3, 1e-3])
    P[:4, 4:] = 0.0
    P[4:, :4] = 0.0
    ids = np.arange(M, dtype=np.int32)
    fb = shim.FilterBatch(shim.UKF_SLAM, p.to_c(), 2, 50, 8)
    fb.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    of.init(0, 0, 0)
    fb.set_state(1, x, P, ids, 7)
    of.set_state(x, P, ids, 7)
    m = np.array([[3, 2.0, 0.3], [11, 1.5, -0.4], [40, 2.5, 0.1]], dtype=np.float32)    # two updates and one insertion
    meas, nm = fb.pack_meas([[], m])
    fb.step(0.05, 0.01, meas, nm)
    of.update(0.05, 0.01, m, oracle.STRUCTURED)
    assert fb.status(1) == 0 and of.status == 0
    assert list(fb.assoc(1)) == list(of.assoc_log()) == [3, 11, -1]
    _compare(fb, 1, of)
    X, Xo = fb.sigma_points(1), of.sigma_points()
    assert X.shape == Xo.shape and H.normwise(X, Xo) <= H.REL_TOL
    assert fb.status(0) == 0 and fb.timestep(0) == 1
