"""Why the reference-faithful UKF-SLAM ends up metres off: a bisection on the NumPy oracle.

The faithful restatement of ukf.cpp diverges to a mean position error of ~2.8 m over 500 steps on the 5x10 grid (the EKF on
the same run: ~0.3 m).  Three reference lines are candidates; each is switched to its "textbook" form in turn:
  * ukf.cpp:310-314  the bearing mean z_est(1) is never accumulated          -> whatif_bearing_mean
  * ukf.cpp:139      the sensing model uses the prior x_t's yaw for every sigma point -> whatif_per_sigma_yaw
  * filter.h:116-117 readCommonParams writes the SENSING covariances into V and leaves W = I -> compat_noise_bug = 0
Result (pinned below): ~80 % of the error follows from filter.h:116-117 alone (W = I tells the filter its range / bearing
measurements have a standard deviation of 1 m / 1 rad, so landmark updates barely correct the pose); the other two lines
change little by themselves.  All three are cited reference behaviour, so the product reproduces them by default
(compat_noise_bug = 1); this test documents the attribution, it does not claim the reference intended it."""
import numpy as np

from oracle import oracle_np
from tests import helpers as H


def _mean_pos_err(p, fwd, ang, stream, truth, compat, bearing_mean=False, per_sigma_yaw=False):
    d = p.as_dict()
    d["landmark_id_is_known"] = 1
    d["compat_noise_bug"] = compat
    f = oracle_np.UKFNP(d)
    f.whatif_bearing_mean, f.whatif_per_sigma_yaw = bearing_mean, per_sigma_yaw
    f.init(0, 0, 0)
    err = []
    for t in range(len(fwd)):
        f.update(fwd[t], ang[t], stream[t])
        err.append(np.hypot(f.x_t[0] - truth[t][0], f.x_t[1] - truth[t][1]))
    return float(np.mean(err)), f.M


def test_ukf_error_is_attributed_to_the_noise_mixup(oracle):
    p, lm, fwd, ang = H.config2(seed=0, steps=500, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    stream, truth = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=0, instance=0)
    faithful, M = _mean_pos_err(p, fwd, ang, stream, truth, 1)
    only_bearing, _ = _mean_pos_err(p, fwd, ang, stream, truth, 1, bearing_mean=True)
    only_yaw, _ = _mean_pos_err(p, fwd, ang, stream, truth, 1, per_sigma_yaw=True)
    only_noise, _ = _mean_pos_err(p, fwd, ang, stream, truth, 0)
    all_three, _ = _mean_pos_err(p, fwd, ang, stream, truth, 0, bearing_mean=True, per_sigma_yaw=True)
    print(f"UKF mean position error over 500 steps (M = {M}): faithful {faithful:.3f} m | bearing mean {only_bearing:.3f} | "
          f"per-sigma yaw {only_yaw:.3f} | corrected noise {only_noise:.3f} | all three {all_three:.3f}")
    assert M >= 20
    assert faithful > 2.0                                    # the reference-faithful filter is metres off
    assert only_bearing > 0.8 * faithful and only_yaw > 0.8 * faithful      # neither line explains it alone
    assert only_noise < 0.3 * faithful                       # filter.h:116-117 carries ~80 % of it
    assert all_three < only_noise < 1.0
    # the C oracle (the parity reference of the CUDA kernels) shows the same two regimes
    for compat, lo, hi in ((1, 2.0, 4.0), (0, 0.2, 1.0)):
        q = H.Params(filter="ukf_slam")
        q.compat_noise_bug = bool(compat)
        f = oracle.OracleFilter(oracle.UKF_SLAM, H.oracle_params(oracle, q), 50)
        f.init(0, 0, 0)
        err = []
        for t in range(len(fwd)):
            f.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
            x = f.state()
            err.append(np.hypot(x[0] - truth[t][0], x[1] - truth[t][1]))
        assert lo < np.mean(err) < hi, (compat, np.mean(err))
