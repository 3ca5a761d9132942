"""The C oracle against the NumPy restatement under NON-default noise parameters (noise means, the corrected-noise branch of
readCommonParams, other covariances): the branches ekf.cpp:57-58,130-131, ukf.cpp:129-131,144-145 and filter.h:105-121 that the
yaml defaults leave at zero / at the V-W mix-up."""
import numpy as np
import pytest

from oracle import oracle_np
from tests import helpers as H


def _np_params(p):
    d = p.as_dict()
    d["landmark_id_is_known"] = int(p.landmark_id_is_known)
    d["compat_noise_bug"] = int(p.compat_noise_bug)
    return d


@pytest.mark.parametrize("variant", sorted(H.PARAM_VARIANTS))
@pytest.mark.parametrize("kind,known", [("ekf", True), ("ekf", False), ("ukf", True)])
def test_c_vs_numpy_param_variants(oracle, variant, kind, known):
    steps = 160 if kind == "ekf" else 70
    p, lm, fwd, ang = H.variant_workload(variant, "ekf_slam" if kind == "ekf" else "ukf_slam", known, seed=3, steps=steps)
    op = H.oracle_params(oracle, p)
    fc = oracle.OracleFilter(oracle.EKF_SLAM if kind == "ekf" else oracle.UKF_SLAM, op, 50)
    fn = (oracle_np.EKFNP if kind == "ekf" else oracle_np.UKFNP)(_np_params(p))
    fc.init(0, 0, 0)
    fn.init(0, 0, 0)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=11, instance=2)
    worst = 0.0
    for t in range(steps):
        fc.update(fwd[t], ang[t], stream[t], oracle.DENSE)
        fn.update(fwd[t], ang[t], stream[t])
        assert list(fc.assoc_log()) == fn.assoc, t
        worst = max(worst, H.normwise(fc.state(), fn.x_t), H.normwise(fc.cov(), fn.P_t))
    assert fc.M == fn.M and list(fc.landmark_ids()) == fn.lm_IDs and fc.M >= 3
    assert worst <= (1e-12 if kind == "ekf" else 1e-10), worst


def test_param_variants_change_the_result(oracle):
    """guard: every variant really moves the estimate away from the default run (the branches are live)"""
    base_p, lm, fwd, ang = H.config2(seed=3, steps=120)
    op0 = H.oracle_params(oracle, base_p)
    stream, _ = H.oracle_meas_stream(oracle, op0, lm, fwd, ang, seed=11, instance=2)
    def run(op, kind):
        f = oracle.OracleFilter(kind, op, 50)
        f.init(0, 0, 0)
        for t in range(len(fwd)):
            f.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
        return f.state()[:3]
    for kind in (oracle.EKF_SLAM, oracle.UKF_SLAM):
        ref = run(op0, kind)
        for name in H.PARAM_VARIANTS:
            x = run(H.oracle_params(oracle, H.variant_params(name)), kind)
            assert np.abs(x - ref).max() > 1e-6, (kind, name)
