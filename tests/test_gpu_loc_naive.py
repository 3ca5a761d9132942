"""GPU parity of the two remaining choices of the reference's `filter:` switch on the hot path: the
localisation-only UKF (FilterChoice::UKF_LOC: true-map sensing model, state stays (x, y, cos, sin)) and the NaiveFilter
(command propagation), through the C-ABI against the CPU oracle."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim():
    from live_ekf_slam_b200 import shim as s
    s.load()
    return s


def test_ukf_loc_batch_vs_oracle(shim, oracle):
    p, lm, fwd, ang = H.config2(seed=7, steps=300, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    B = 6
    fb = shim.FilterBatch(shim.UKF_LOC, p.to_c(), B, 50, 8)
    fb.set_map(lm)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=8, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.UKF_LOC, op, 50)
        of.init(0, 0, 0)
        of.set_map(lm)
        ofs.append(of)
    worst, n_upd = 0.0, 0
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([streams[i][t] for i in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], streams[i][t], oracle.DENSE)
        n_upd += int(n.sum())
        if t % 10 == 0 or t > 290:
            for i in range(B):
                assert list(fb.assoc(i)) == list(ofs[i].assoc_log()), (t, i)
                assert fb.num_landmarks(i) == 0
                worst = max(worst, H.normwise(fb.state(i), ofs[i].state()), H.normwise(fb.cov(i), ofs[i].cov()))
    assert worst <= H.REL_TOL and n_upd > 500, (worst, n_upd)
    poses = fb.poses()
    for i in range(B):
        xo = ofs[i].state()
        assert np.abs(poses[i, :2] - xo[:2]).max() <= H.FINAL_TOL and abs(poses[i, 2] - np.arctan2(xo[3], xo[2])) <= H.FINAL_TOL
    assert (fb.all_status() == 0).all()
    print("ukf_loc worst normwise err", worst)


def test_ukf_loc_bad_id_and_map_guard(shim, oracle):
    p = H.Params(filter="ukf_slam")
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(shim.UKF_LOC, p.to_c(), 2, 4, 3)
    lm = np.array([[1.0, 0.5], [2.0, -0.5]])
    fb.set_map(lm)
    fb.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_LOC, op, 4)
    of.init(0, 0, 0)
    of.set_map(lm)
    m = np.array([[1, 2.0, -0.2], [7, 1.0, 0.1]], dtype=np.float32)        # id 7 is outside the map
    meas, n = fb.pack_meas([m, []])
    fb.step(0.05, 0.0, meas, n)
    of.update(0.05, 0.0, m)
    assert fb.status(0) & shim.STATUS_BAD_ID and of.status & oracle.ERR_BAD_ID and fb.status(1) == 0
    assert list(fb.assoc(0)) == [1, -1] == list(of.assoc_log())
    assert H.normwise(fb.state(0), of.state()) <= H.REL_TOL and H.normwise(fb.cov(0), of.cov()) <= H.REL_TOL
    ek = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 4, 3)
    with pytest.raises(shim.SlamError):                                   # only the localisation-only UKF keeps a map
        ek.set_map(lm)


def test_ukf_loc_filter_switch(shim, oracle):
    """`filter: ukf_loc` through the Python mirror of the plugin interface (localization_node.cpp:36-38)."""
    from live_ekf_slam_b200.filter import make_filter, UKF, FilterChoice
    p, lm, fwd, ang = H.config2(seed=9, steps=60, filt="ukf_slam")
    p.filter = "ukf_loc"
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=4, instance=0)
    filt = make_filter(p, max_landmarks=50, max_meas=8)
    assert isinstance(filt, UKF) and filt.type == FilterChoice.UKF_LOC
    wire = np.zeros((len(lm), 3), dtype=np.float32)
    wire[:, 0] = np.arange(len(lm)); wire[:, 1:] = lm
    filt.setMap(wire.reshape(-1))
    filt.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_LOC, op, 50)
    of.init(0, 0, 0)
    of.set_map(lm)
    for t in range(len(fwd)):
        filt.update((fwd[t], ang[t]), stream[t].reshape(-1))
        of.update(fwd[t], ang[t], stream[t])
    xo = of.state()
    sv = filt.getStateVector()
    assert sv.size == 3 and abs(sv[0] - xo[0]) <= H.REL_TOL and abs(sv[2] - np.arctan2(xo[3], xo[2])) <= 1e-9
    msg = filt.publishState()
    assert msg["M"] == 0 and msg["P"].size == 16 and filt.timestep == len(fwd)


def test_naive_filter_vs_oracle(shim, oracle):
    from live_ekf_slam_b200.filter import NaiveFilter
    p, lm, fwd, ang = H.config2(seed=2, steps=150)
    op = H.oracle_params(oracle, p)
    B = 3
    fb = shim.FilterBatch(shim.NAIVE, p.to_c(), B, 1, 4)
    fb.init(0.5, -1.0, 0.25)
    of = oracle.OracleFilter(oracle.NAIVE, op, 1)
    of.init(0.5, -1.0, 0.25)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([[] for _ in range(B)])
        fb.step(fwd[t], ang[t], meas, n)
        of.update(fwd[t], ang[t], [])
    for i in range(B):
        assert np.abs(fb.state(i) - of.state()).max() <= 1e-12 and fb.timestep(i) == len(fwd)
    nf = NaiveFilter(max_landmarks=1, max_meas=4)
    nf.readParams(p)
    nf.init(0.5, -1.0, 0.25)
    for t in range(20):
        nf.update((fwd[t], ang[t]), [3.0, 1.0, 0.0])
    msg = nf.publishState()
    assert msg["timestep"] == 20 and set(msg) == {"timestep", "x_v", "y_v", "yaw_v"}
