"""Independent NumPy restatement of the reference filters.  TEST INFRASTRUCTURE ONLY.

Second, independently written restatement of ekf.cpp:37-179 and ukf.cpp:106-371 (LAPACK eigh, BLAS
products, np.float32 scalars for the reference's float roundings).  Its only job is to cross-check
oracle/slam_oracle.c so that a single mis-transcription cannot silently define "truth"
(SURVEY.md section 4 / 8c).  PARITY UNPINNED (no reference tests or golden vectors exist).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
PI = 3.14159265358979323846  # filter.h:42
TWO_PI = 2 * PI


def _cosf(x):  # D-1: cos(float) pinned as (float)cos((double)x)
    return F(math.cos(float(x)))


def _sinf(x):
    return F(math.sin(float(x)))


class FilterNP:
    """Shared fields + readCommonParams, filter.h:79-121."""

    def __init__(self, params: dict):
        p = params
        self.v_d, self.v_th = F(p["v_d"]), F(p["v_th"])
        self.w_r, self.w_b = F(p["w_r"]), F(p["w_b"])
        self.V = np.diag([p["V_00"], p["V_11"]]).astype(float)
        self.W = np.eye(2)
        if p.get("compat_noise_bug", 1):
            self.V = np.diag([p["W_00"], p["W_11"]]).astype(float)  # filter.h:116-117
        else:
            self.W = np.diag([p["W_00"], p["W_11"]]).astype(float)
        self.known = bool(p["landmark_id_is_known"])
        self.sep = F(p["min_landmark_separation"])
        self.timestep = 0
        self.M = 0
        self.lm_IDs: list[int] = []
        self.assoc: list[int] = []


class EKFNP(FilterNP):
    def __init__(self, params):
        super().__init__(params)
        self.x_t = np.zeros(3)
        self.P_t = np.diag([0.01 * 0.01, 0.01 * 0.01, 0.005 * 0.005])  # ekf.cpp:11-14

    def init(self, x0, y0, yaw0):
        self.x_t = np.array([float(F(x0)), float(F(y0)), float(F(yaw0))])  # ekf.cpp:31

    def update(self, fwd, ang, meas):
        self.timestep += 1
        d_d, d_th = F(fwd), F(ang)
        n = 3 + 2 * self.M
        th = self.x_t[2]
        F_x = np.eye(n)
        F_x[0, 2] = float(F(-1) * d_d) * math.sin(th)
        F_x[1, 2] = float(d_d) * math.cos(th)
        F_v = np.zeros((n, 2))
        F_v[0, 0] = math.cos(th); F_v[1, 0] = math.sin(th); F_v[2, 1] = 1
        x_pred = self.x_t.copy()
        dv = float(F(d_d + self.v_d))
        x_pred[0] = self.x_t[0] + dv * math.cos(th)
        x_pred[1] = self.x_t[1] + dv * math.sin(th)
        x_pred[2] = math.remainder(self.x_t[2] + float(d_th) + float(self.v_th), TWO_PI)
        P_pred = F_x @ self.P_t @ F_x.T + F_v @ self.V @ F_v.T
        meas = np.asarray(meas, dtype=np.float32).reshape(-1, 3)
        self.assoc = []
        x_stale = self.x_t  # ekf.cpp:115 reads landmark means from x_t
        for l in range(len(meas)):
            r, b = F(meas[l, 1]), F(meas[l, 2])
            i = -1
            if not self.known:
                ident = self.M
                xd = F(x_pred[0] + float(r) * math.cos(x_pred[2] + float(b)))
                yd = F(x_pred[1] + float(r) * math.sin(x_pred[2] + float(b)))
                for j in range(self.M):
                    x_diff = F(abs(float(xd) - x_pred[3 + 2 * j]))
                    y_diff = F(abs(float(yd) - x_pred[4 + 2 * j]))
                    if x_diff < self.sep and y_diff < self.sep:
                        i = j; ident = j
                        break
            else:
                ident = int(meas[l, 0])
                for j in range(self.M):
                    if self.lm_IDs[j] == ident:
                        i = j
                        break
            self.assoc.append(i)
            n = 3 + 2 * self.M
            if i != -1:
                i = i * 2 + 3
                if i + 1 >= len(x_stale):
                    raise RuntimeError("same-step re-match (reference indexes x_t out of range)")
                dx = x_stale[i] - x_pred[0]
                dy = x_stale[i + 1] - x_pred[1]
                dist = F(math.sqrt(dx * dx + dy * dy))
                dd = float(dist)
                d2 = float(F(dist * dist))
                H = np.zeros((2, n))
                H[0, 0] = -dx / dd; H[0, 1] = -dy / dd
                H[1, 0] = dy / d2; H[1, 1] = -dx / d2; H[1, 2] = -1
                H[0, i] = dx / dd; H[0, i + 1] = dy / dd
                H[1, i] = -dy / d2; H[1, i + 1] = dx / d2
                angf = F(math.remainder(math.atan2(dy, dx) - x_pred[2], TWO_PI))
                nu = np.array([float(F(F(r - dist) - self.w_r)), float(F(F(b - angf) - self.w_b))])
                S = H @ P_pred @ H.T + self.W
                K = P_pred @ H.T @ np.linalg.inv(S)
                x_pred = x_pred + K @ nu
                x_pred[2] = math.remainder(x_pred[2], TWO_PI)
                P_pred = P_pred - (K @ H) @ P_pred
            else:
                self.M += 1
                nn = 3 + 2 * self.M
                a = x_pred[2] + float(b)
                cb, sb = math.cos(a), math.sin(a)
                x_pred = np.concatenate([x_pred, [x_pred[0] + float(r) * cb, x_pred[1] + float(r) * sb]])
                self.lm_IDs.append(ident)
                Y = np.eye(nn)
                Y[nn - 2, nn - 2] = cb; Y[nn - 2, nn - 1] = -float(r) * sb
                Y[nn - 1, nn - 2] = sb; Y[nn - 1, nn - 1] = float(r) * cb
                Y[nn - 2, 0] = 1; Y[nn - 2, 2] = -float(r) * sb
                Y[nn - 1, 1] = 1; Y[nn - 1, 2] = float(r) * cb
                pt = np.zeros((nn, nn))
                pt[: nn - 2, : nn - 2] = P_pred
                pt[nn - 2:, nn - 2:] = self.W
                P_pred = Y @ pt @ Y.T
        self.x_t = x_pred
        self.P_t = P_pred


class UKFNP(FilterNP):
    def __init__(self, params):
        super().__init__(params)
        self.W_0 = F(0.2)  # filter.h:207
        self.x_t = np.zeros(4)
        self.P_t = np.diag([1e-2 * 1e-2, 1e-2 * 1e-2, 0.005 * 0.005, 0.005 * 0.005])
        self.X = None
        self.X_pred = None
        self.loc = False        # FilterChoice::UKF_LOC (localization_node.cpp:36-38): landmarks from the true map
        self.map = None         # float32 [id, x, y]*, filter.h:68
        # WHAT-IF switches, off by default (= the reference).  Only tests/test_ukf_divergence_bisection.py turns them on, to
        # attribute the UKF's large position error to individual reference lines:
        self.whatif_bearing_mean = False    # accumulate z_est(1) too (the reference never does, ukf.cpp:310-314)
        self.whatif_per_sigma_yaw = False   # sensing model with each sigma point's own yaw (the reference uses x_t's, ukf.cpp:139)

    def set_map(self, lm_xy):
        lm = np.asarray(lm_xy, dtype=np.float64).reshape(-1, 2)
        m = np.zeros((len(lm), 3), dtype=np.float32)
        m[:, 0] = np.arange(len(lm)); m[:, 1:] = lm.astype(np.float32)
        self.map = m.reshape(-1)

    @staticmethod
    def _yaw(x):
        return F(math.remainder(math.atan2(x[3], x[2]), TWO_PI))

    def init(self, x0, y0, yaw0):
        y = F(yaw0)
        self.x_t = np.array([float(F(x0)), float(F(y0)), float(_cosf(y)), float(_sinf(y))])

    def _weights(self, n):
        w = F(F(1 - self.W_0) / F(2 * n))
        Wts = np.full(2 * n + 1, float(w))
        Wts[0] = float(self.W_0)
        return Wts

    def _motion(self, x, u_d, u_th):
        out = x.copy()
        yaw = self._yaw(x)
        ud = F(u_d + self.v_d)
        out[0] = x[0] + float(F(ud * _cosf(yaw)))
        out[1] = x[1] + float(F(ud * _sinf(yaw)))
        fs = F(F(yaw + u_th) + self.v_th)
        new_yaw = F(math.remainder(float(fs), TWO_PI))
        out[2] = float(_cosf(new_yaw))
        out[3] = float(_sinf(new_yaw))
        return out

    def update(self, fwd, ang, meas):
        self.timestep += 1
        u_d, u_th = F(fwd), F(ang)
        n = 4 + 2 * self.M
        Wts = self._weights(n)
        yaw = self._yaw(self.x_t)
        Q = np.zeros((n, n))
        Q[0, 0] = self.V[0, 0] * float(_cosf(yaw)); Q[1, 1] = self.V[0, 0] * float(_sinf(yaw))
        Q[2, 2] = self.V[1, 1] * float(_cosf(yaw)); Q[3, 3] = self.V[1, 1] * float(_sinf(yaw))
        # nearestSPD + sqrt, ukf.cpp:106-123,208
        scale = float(F(F(2 * self.M + 4) / F(1 - self.W_0)))
        Y = 0.5 * (self.P_t + self.P_t.T) * scale
        D, Qv = np.linalg.eigh(Y)
        Dp = np.maximum(D, 0.00000001)
        Yp = (Qv * Dp) @ Qv.T
        lam, U = np.linalg.eigh(0.5 * (Yp + Yp.T))
        sq = (U * np.sqrt(np.maximum(lam, 0.0))) @ U.T
        X = np.zeros((n, 2 * n + 1))
        X[:, 0] = self.x_t
        for i in range(1, n + 1):
            X[:, i] = self.x_t + sq[:, i - 1]
            X[:, i + n] = self.x_t - sq[:, i - 1]
        Xp = np.stack([self._motion(X[:, i], u_d, u_th) for i in range(2 * n + 1)], axis=1)
        x_pred = np.zeros(n)
        for i in range(2 * n + 1):
            x_pred = x_pred + Wts[i] * Xp[:, i]
        Dm = Xp - x_pred[:, None]
        P_pred = (Dm * Wts) @ Dm.T + Q
        self.X, self.X_pred = X, Xp
        # updateStage
        meas = np.asarray(meas, dtype=np.float32).reshape(-1, 3)
        self.assoc = []
        new = []
        yaw_prior = self._yaw(self.x_t)
        for l in range(len(meas)):
            ident = int(meas[l, 0]); r = F(meas[l, 1]); b = F(meas[l, 2])
            lm_i = -1
            if self.loc:                                     # ukf.cpp:262,272,300-302
                lm_i = ident
                self.assoc.append(lm_i)
                dx = float(self.map[lm_i * 3 + 1]) - Xp[0, :]    # :152-153
                dy = float(self.map[lm_i * 3 + 2]) - Xp[1, :]
            else:
                for j in range(self.M):
                    if self.lm_IDs[j] == ident:
                        lm_i = j
                        break
                self.assoc.append(lm_i)
                if lm_i == -1:
                    new.append(l)
                    continue
                li = lm_i * 2 + 4
                dx = Xp[li, :] - Xp[0, :]
                dy = Xp[li + 1, :] - Xp[1, :]
            z0 = np.sqrt(dx * dx + dy * dy) + float(self.w_r)
            yaws = ([float(self._yaw(Xp[:, c])) for c in range(2 * n + 1)] if self.whatif_per_sigma_yaw
                    else [float(yaw_prior)] * (2 * n + 1))
            z1 = np.array([math.remainder(math.atan2(dy[c], dx[c]) - yaws[c] + float(self.w_b), TWO_PI)
                           for c in range(2 * n + 1)])
            zest0 = 0.0
            for c in range(2 * n + 1):
                zest0 += Wts[c] * z0[c]
            zest1 = float(np.dot(Wts, z1)) if self.whatif_bearing_mean else 0.0
            d0 = z0 - zest0
            d1 = np.array([math.remainder(v - zest1, TWO_PI) for v in z1])
            Dz = np.stack([d0, d1])
            S = (Dz * Wts) @ Dz.T + self.W
            Cm = ((Xp - x_pred[:, None]) * Wts) @ Dz.T
            K = Cm @ np.linalg.inv(S)
            innov = np.array([float(r) - zest0, math.remainder(float(b) - zest1, TWO_PI)])
            x_pred = x_pred + K @ innov
            P_pred = P_pred - K @ S @ K.T
        for l in new:
            ident = int(meas[l, 0]); r = F(meas[l, 1]); b = F(meas[l, 2])
            nn = 4 + 2 * self.M
            yawp = self._yaw(x_pred)
            yb = F(yawp + b)
            x_pred = np.concatenate([x_pred, [x_pred[0] + float(F(r * _cosf(yb))), x_pred[1] + float(F(r * _sinf(yb)))]])
            self.lm_IDs.append(ident)
            pt = np.eye(nn + 2)
            pt[:nn, :nn] = P_pred
            pt[nn:, nn:] = self.W
            P_pred = pt
            self.M += 1
        self.x_t = x_pred
        self.P_t = P_pred


def sim_step_np(p: dict, truth, fwd, ang, lm_xy, u_cmd, u_lm):
    """sim_node.py:209-250 with the uniforms supplied by the caller:
    u_cmd = (u_fwd, u_ang); u_lm[id] = (u_r, u_b).  Returns (new_truth, float32 [k,3])."""
    d = float(F(fwd)) + 2 * p["V_00"] * u_cmd[0] - p["V_00"]
    hdg = float(F(ang)) + 2 * p["V_11"] * u_cmd[1] - p["V_11"]
    d = max(0, min(d, p["d_max"]))
    hdg = max(-p["th_max"], min(hdg, p["th_max"]))
    x_v = [truth[0] + d * math.cos(truth[2]), truth[1] + d * math.sin(truth[2]), truth[2] + hdg]
    out = []
    for ident in range(len(lm_xy)):
        dx = lm_xy[ident][0] - x_v[0]
        dy = lm_xy[ident][1] - x_v[1]
        r = math.sqrt(dx * dx + dy * dy)
        beta = math.remainder(math.atan2(dy, dx) - x_v[2], math.tau)
        if r > p["range_max"]:
            continue
        elif beta > p["fov_min"] and beta < p["fov_max"]:
            ur, ub = u_lm[ident]
            out.append([ident, r + 2 * p["W_00"] * ur - p["W_00"], beta + 2 * p["W_11"] * ub - p["W_11"]])
    return x_v, np.asarray(out, dtype=np.float32).reshape(-1, 3)
