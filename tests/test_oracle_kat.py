"""Known-answer tests for the CPU oracle, hand-derived from the cited reference lines.
The reference ships no tests or golden vectors (PARITY UNPINNED); these pin the restatement to the equations."""
import numpy as np
import pytest

from oracle import oracle_np


def test_kat1_ekf_predict_no_detection(oracle):
    # ekf.cpp:47-61 with init(0,0,0), cmd (0.1f, 0), V = diag(.01,.01) (noise bug), no detection
    for mode in (oracle.DENSE, oracle.STRUCTURED):
        f = oracle.OracleFilter(oracle.EKF_SLAM)
        f.init(0, 0, 0)
        f.update(0.1, 0.0, [], mode)
        d = float(np.float32(0.1))
        np.testing.assert_array_equal(f.state(), [d, 0.0, 0.0])
        P = f.cov()
        exp = np.array([[1e-4 + 0.01, 0, 0],
                        [0, 1e-4 + d * d * 2.5e-5, d * 2.5e-5],
                        [0, d * 2.5e-5, 2.5e-5 + 0.01]])
        np.testing.assert_allclose(P, exp, rtol=1e-15, atol=1e-20)
        assert f.timestep == 1 and f.M == 0


def test_kat2_ekf_first_insertion(oracle):
    # ekf.cpp:144-172 same step with meas [7, 1.0, 0.0]
    f = oracle.OracleFilter(oracle.EKF_SLAM)
    f.init(0, 0, 0)
    f.update(0.1, 0.0, [7, 1.0, 0.0], oracle.DENSE)
    d = float(np.float32(0.1))
    np.testing.assert_allclose(f.state(), [d, 0, 0, d + 1.0, 0.0], rtol=1e-15)
    assert list(f.landmark_ids()) == [7] and f.M == 1
    P = f.cov()
    np.testing.assert_allclose(P[3:, 3:], [[1.0101, 0], [0, 1.0025e-4 + 2 * 2.5e-6 + 0.010025 + 1]], rtol=1e-7)
    np.testing.assert_allclose(P[3:, 3:], [[1.0101, 0], [0, 1.01013025]], rtol=1e-8)
    # cross block rows = G_x * P[0:3, :]
    Pv = P[:3, :3]
    Gx = np.array([[1, 0, 0], [0, 1, 1.0]])
    np.testing.assert_allclose(P[3:, :3], Gx @ Pv, rtol=1e-14, atol=1e-20)
    np.testing.assert_allclose(P[:3, 3:], Pv @ Gx.T, rtol=1e-14, atol=1e-20)


def test_kat3_ukf_weights():
    # ukf.cpp:35,114,175 float arithmetic (values measured in SURVEY App. A)
    u = oracle_np.UKFNP(dict(v_d=0, v_th=0, w_r=0, w_b=0, V_00=.01, V_11=.001, W_00=.01, W_11=.01,
                             landmark_id_is_known=1, min_landmark_separation=.1))
    w4 = u._weights(4)
    assert w4[1] == float(np.float32(0.1)) and w4[0] == float(np.float32(0.2))
    assert abs(w4.sum() - 1 - 1.49e-8) < 1e-10
    w104 = u._weights(104)
    assert abs(w104[1] - 0.0038461538497358561) < 1e-18
    assert abs(w104.sum() - 1 - 3.73e-9) < 2e-11
    assert float(np.float32(104) / np.float32(1 - np.float32(0.2))) == 130.0


def test_kat3_ukf_weights_in_c_oracle(oracle):
    # after one no-detection step from P0 the state mean of rows 0,1 equals sum(w)*x + motion: check sum(w) != 1
    f = oracle.OracleFilter(oracle.UKF_SLAM)
    f.init(1.0, 2.0, 0.0)
    f.update(0.0, 0.0, [], oracle.DENSE)
    x = f.state()
    sw = 0.2 * 0 + float(np.float32(0.2)) + 8 * float(np.float32(0.1))
    assert abs(x[0] - sw * 1.0) < 1e-12 and abs(x[1] - sw * 2.0) < 1e-12
    assert abs(sw - 1 - 1.49e-8) < 1e-10


def test_kat4_ukf_bearing_mean_is_zero(oracle):
    # ukf.cpp:310-314 only z_est(0) is accumulated; the bearing innovation is therefore the raw bearing.
    # Two filters that differ only in the measured bearing by 2*pi*k must agree; a filter fed bearing b gets
    # innovation(1) = remainder(b, 2pi) regardless of where the landmark is.
    pd = dict(v_d=0, v_th=0, w_r=0, w_b=0, V_00=.01, V_11=.001, W_00=.01, W_11=.01, landmark_id_is_known=1,
              min_landmark_separation=.1)
    u = oracle_np.UKFNP(pd)
    u.init(0, 0, 0)
    u.update(0.1, 0.0, [3, 2.0, 0.5])
    c = oracle.OracleFilter(oracle.UKF_SLAM)
    c.init(0, 0, 0)
    c.update(0.1, 0.0, [3, 2.0, 0.5])
    x_before = c.state().copy()
    # second observation with bearing 0: innovation(1) == 0 exactly, so only the range residual moves the state
    c.update(0.0, 0.0, [3, 2.0, 0.0])
    u.update(0.0, 0.0, [3, 2.0, 0.0])
    assert np.abs(c.state() - u.x_t).max() < 1e-12
    assert c.M == 1 and x_before.shape == (6,)


def test_kat5_ukf_landmark_block_property(oracle):
    # after any UKF predict the landmark block of P_pred equals (2 w scale) * clipSPD(sym(P))[landmark block]
    # because the motion model leaves rows >= 4 untouched (ukf.cpp:127-133).  Relative 1e-7.
    f = oracle.OracleFilter(oracle.UKF_SLAM)
    f.init(0, 0, 0)
    f.update(0.1, 0.01, [1, 2.0, 0.3, 4, 1.5, -0.4])   # two insertions
    P0 = f.cov()
    n = f.n
    f.update(0.1, 0.01, [])                              # predict only
    P1 = f.cov()
    w = float(np.float32(np.float32(1 - np.float32(0.2)) / np.float32(2 * n)))
    scale = float(np.float32(np.float32(n) / np.float32(1 - np.float32(0.2))))
    Y = 0.5 * (P0 + P0.T) * scale
    d, Q = np.linalg.eigh(Y)
    Yp = (Q * np.maximum(d, 1e-8)) @ Q.T
    np.testing.assert_allclose(P1[4:, 4:], 2 * w * Yp[4:, 4:], rtol=1e-7, atol=1e-12)


def test_ukf_first_step_Q(oracle):
    # Q = diag(.01 cosf(0), .01 sinf(0), .01 cosf(0), .01 sinf(0)) = diag(.01, 0, .01, 0) on the first step
    f = oracle.OracleFilter(oracle.UKF_SLAM)
    f.init(0, 0, 0)
    f.update(0.0, 0.0, [])
    P = f.cov()
    # with no motion the sigma spread reproduces ~P0 (times 2 w scale) and Q adds on the diagonal
    assert abs(P[0, 0] - (1e-4 + 0.01)) < 1e-9
    assert abs(P[1, 1] - 1e-4) < 1e-9


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    F = 0xffffffff
    assert oracle.philox(F, F, F, F, F, F) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    assert oracle.uniform(0, 0) == 0.0
    assert oracle.uniform(F, F) == 1.0 - 2.0 ** -53


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 43, 104])
def test_eigh_against_lapack(oracle, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = A + A.T
    w, V = oracle.eigh(A)
    w2 = np.linalg.eigvalsh(A)
    scale = max(1.0, np.abs(w2).max())
    assert np.abs(w - w2).max() < 1e-13 * n * scale
    assert np.abs(V @ np.diag(w) @ V.T - A).max() < 1e-13 * n * scale
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-13 * n
