"""GPU parity under NON-default parameters: noise means v_d, v_th, w_r, w_b != 0 (the float adds of
ekf.cpp:57-58,130-131 and ukf.cpp:129-131,144-145), the corrected-noise branch (compat_noise_bug = 0: V and W as the yaml
states them, filter.h:105-121 without the mix-up of :116-117), other covariances, unknown-ID association -- for the per-step
kernels, the persistent sweep kernel and the simulator.  Tolerances as everywhere: decisions bit-exact, state and covariance
1e-9 norm-wise."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

VARIANTS = sorted(H.PARAM_VARIANTS)


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


def _free_run(shim, oracle, skind, okind, p, lm, fwd, ang, B, seed, every=20):
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(skind, p.to_c(), B, 50, 8)
    fb.init(0, 0, 0)
    streams = [H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=seed, instance=i)[0] for i in range(B)]
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(okind, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    worst, matched = 0.0, 0
    for t in range(len(fwd)):
        msgs = [streams[i][t].copy() for i in range(B)]
        if not p.landmark_id_is_known:
            for m in msgs:
                if len(m):
                    m[:, 0] = 777.0                      # ids on the wire are ignored in this mode
        meas, n = fb.pack_meas(msgs)
        fb.step(fwd[t], ang[t], meas, n)
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], msgs[i], oracle.STRUCTURED)
        if t % every == 0 or t == len(fwd) - 1:
            for i in range(B):
                a = list(fb.assoc(i))
                assert a == list(ofs[i].assoc_log()), (t, i)
                matched += sum(1 for v in a if v >= 0)
                worst = max(worst, _compare(fb, i, ofs[i]))
    assert (fb.all_status() == 0).all() and all(o.status == 0 for o in ofs)
    return worst, matched, ofs, fb


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("known", [True, False], ids=["known_ids", "box_gate"])
def test_ekf_param_variants(shim, oracle, variant, known):
    p, lm, fwd, ang = H.variant_workload(variant, "ekf_slam", known, seed=4, steps=300)
    worst, matched, ofs, fb = _free_run(shim, oracle, shim.EKF_SLAM, oracle.EKF_SLAM, p, lm, fwd, ang, B=6, seed=21)
    assert ofs[0].M >= 8 and matched > 20
    x, xo = fb.state(0), ofs[0].state()
    assert np.abs(x[:3] - xo[:3]).max() <= H.FINAL_TOL
    print("ekf", variant, "known" if known else "box gate", "worst", worst)


@pytest.mark.parametrize("variant", VARIANTS)
def test_ukf_param_variants(shim, oracle, variant):
    p, lm, fwd, ang = H.variant_workload(variant, "ukf_slam", True, seed=5, steps=200)
    worst, matched, ofs, fb = _free_run(shim, oracle, shim.UKF_SLAM, oracle.UKF_SLAM, p, lm, fwd, ang, B=4, seed=22, every=10)
    assert ofs[0].M >= 6 and matched > 10
    print("ukf", variant, "worst", worst)


@pytest.mark.parametrize("variant", VARIANTS)
def test_ukf_loc_param_variants(shim, oracle, variant):
    p, lm, fwd, ang = H.variant_workload(variant, "ukf_slam", True, seed=6, steps=150)
    op = H.oracle_params(oracle, p)
    fb = shim.FilterBatch(shim.UKF_LOC, p.to_c(), 2, 50, 8)
    fb.set_map(lm)
    fb.init(0, 0, 0)
    of = oracle.OracleFilter(oracle.UKF_LOC, op, 50)
    of.init(0, 0, 0)
    of.set_map(lm)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=23, instance=1)
    for t in range(len(fwd)):
        meas, n = fb.pack_meas([stream[t], stream[t]])
        fb.step(fwd[t], ang[t], meas, n)
        of.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
        if t % 10 == 0:
            assert H.normwise(fb.state(1), of.state()) <= H.REL_TOL and H.normwise(fb.cov(1), of.cov()) <= H.REL_TOL, t
    assert H.normwise(fb.state(0), of.state()) <= H.REL_TOL


@pytest.mark.parametrize("variant", VARIANTS)
def test_sweep_kernel_param_variants(shim, oracle, variant):
    """slam_run (persistent sweep kernel: on-device simulator -> filter) under the variant, against the oracle fed the
    DEVICE's own messages (recorded step by step from a twin handle), so the comparison stays at 1e-9 per step."""
    p, lm, fwd, ang = H.variant_workload(variant, "ekf_slam", True, seed=7, steps=220)
    op = H.oracle_params(oracle, p)
    B, seed, off = 10, 99, 40
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    fb.init(0, 0, 0)
    sim = shim.Simulator(fb, lm, seed=seed, instance_offset=off)
    sim.run(fwd, ang)
    fb.synchronize()
    # twin: the same simulator stepped alone, its messages read back and fed to the oracle
    tw = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
    tsim = shim.Simulator(tw, lm, seed=seed, instance_offset=off)
    ofs = []
    for i in range(B):
        of = oracle.OracleFilter(oracle.EKF_SLAM, op, 50)
        of.init(0, 0, 0)
        ofs.append(of)
    for t in range(len(fwd)):
        tsim.step(fwd[t], ang[t], t)
        m, n = tsim.meas()
        for i in range(B):
            ofs[i].update(fwd[t], ang[t], m[i, : n[i]], oracle.STRUCTURED)
    np.testing.assert_array_equal(sim.truth(), tsim.truth())
    for i in range(B):
        _compare(fb, i, ofs[i])
        assert fb.timestep(i) == len(fwd)


@pytest.mark.parametrize("sim_kw", [dict(V_00=0.03, V_11=0.004, W_00=0.05, W_11=0.02),
                                    dict(d_max=0.07, th_max=0.03, range_max=4.5, fov_min=-0.9, fov_max=1.2)],
                         ids=["noise_half_widths", "constraints"])
def test_sim_param_variants(shim, oracle, sim_kw):
    """sim_node.py:209-250 under other noise half-widths (:216-217,247-248) and other command / vision constraints
    (:219-220,239-241): ids, visibility decisions and truth exact, float32 r / b within one float ulp."""
    p, lm, fwd, ang = H.config2(seed=1, steps=180)
    for k, v in sim_kw.items():
        setattr(p, k, v)
    op = H.oracle_params(oracle, p)
    B, seed, off = 8, 31337, 11
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 16)
    sim = shim.Simulator(fb, lm, seed=seed, instance_offset=off)
    truths = [np.zeros(3) for _ in range(B)]
    n_msgs, flips = 0, 0
    for t in range(len(fwd)):
        sim.step(fwd[t], ang[t], t)
        m, n = sim.meas()
        tr = sim.truth()
        for i in range(B):
            ref = oracle.sim_step(op, truths[i], fwd[t], ang[t], lm, seed, off + i, t)
            assert n[i] == len(ref), (t, i)
            got = m[i, : n[i]]
            np.testing.assert_array_equal(got[:, 0], ref[:, 0])
            if not np.array_equal(got, ref):
                np.testing.assert_allclose(got, ref, rtol=1.3e-7, atol=0)
                flips += int((got != ref).sum())
            n_msgs += len(ref)
            assert np.abs(tr[i] - truths[i]).max() <= 1e-12
    assert n_msgs > 1000 and flips <= 2, (n_msgs, flips)
