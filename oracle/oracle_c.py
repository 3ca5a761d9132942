"""ctypes binding of the C oracle (oracle/slam_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (live_ekf_slam_b200/) never does.
PARITY UNPINNED: see oracle/slam_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libslam_oracle.so")

EKF_SLAM = 1
UKF_LOC = 2
UKF_SLAM = 3
NAIVE = 5
DENSE = 0
STRUCTURED = 1
ERR_NAN, ERR_SAME_STEP_REMATCH, ERR_CAPACITY = 1, 2, 4
ERR_BAD_ID = 16


class OracleParams(C.Structure):
    _fields_ = [
        ("v_d", C.c_float), ("v_th", C.c_float), ("w_r", C.c_float), ("w_b", C.c_float),
        ("V_00", C.c_double), ("V_11", C.c_double), ("W_00", C.c_double), ("W_11", C.c_double),
        ("landmark_id_is_known", C.c_int), ("min_landmark_separation", C.c_float),
        ("compat_noise_bug", C.c_int),
        ("d_max", C.c_double), ("th_max", C.c_double), ("range_max", C.c_double),
        ("fov_min", C.c_double), ("fov_max", C.c_double),
    ]


#: BP/config/params.yaml defaults (lines 27-52)
DEFAULTS = dict(v_d=0.0, v_th=0.0, w_r=0.0, w_b=0.0, V_00=0.01, V_11=0.001, W_00=0.01, W_11=0.01,
                landmark_id_is_known=1, min_landmark_separation=0.1, compat_noise_bug=1,
                d_max=0.1, th_max=0.0546, range_max=3.0, fov_min=-1.57, fov_max=1.57)


def make_params(**kw) -> OracleParams:
    d = dict(DEFAULTS)
    d.update(kw)
    return OracleParams(**d)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "slam_oracle.c")
    hdr = os.path.join(_HERE, "slam_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libslam_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)
    pp = C.POINTER(OracleParams)
    L.oracle_create.restype = C.c_void_p
    L.oracle_create.argtypes = [C.c_int, pp, C.c_int]
    L.oracle_destroy.argtypes = [C.c_void_p]
    L.oracle_init.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.oracle_set_map.argtypes = [C.c_void_p, fp, C.c_int]
    L.oracle_update.argtypes = [C.c_void_p, C.c_float, C.c_float, fp, C.c_int, C.c_int]
    L.oracle_predict.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
    L.oracle_measure.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
    L.oracle_set_state.argtypes = [C.c_void_p, dp, dp, ip, C.c_int, C.c_int]
    for name in ("oracle_state_dim", "oracle_num_landmarks", "oracle_timestep", "oracle_status"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.oracle_get_state.argtypes = [C.c_void_p, dp]
    L.oracle_get_cov.argtypes = [C.c_void_p, dp]
    L.oracle_get_landmark_ids.argtypes = [C.c_void_p, ip]
    L.oracle_get_assoc_log.argtypes = [C.c_void_p, ip, C.c_int]
    L.oracle_get_sigma_points.argtypes = [C.c_void_p, dp]
    L.oracle_sigma_rows.argtypes = [C.c_void_p]
    L.oracle_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    L.oracle_uniform.restype = C.c_double
    L.oracle_uniform.argtypes = [C.c_uint32, C.c_uint32]
    L.oracle_sim_step.argtypes = [pp, dp, C.c_float, C.c_float, dp, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, fp, C.c_int]
    L.oracle_run_instance.argtypes = [C.c_int, pp, dp, C.c_int, fp, fp, C.c_int, C.c_uint64, C.c_uint32,
                                      C.c_int, C.c_int, dp, dp, C.POINTER(C.c_void_p)]
    L.oracle_bench.restype = C.c_double
    L.oracle_bench.argtypes = [C.c_int, pp, dp, C.c_int, fp, fp, C.c_int, C.c_uint64, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    L.oracle_eigh.argtypes = [dp, C.c_int, dp, dp]
    L.oracle_tsp_trajectory.argtypes = [pp, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_int, C.c_uint64, C.c_uint32, fp, fp]
    L.oracle_set_trig_mode.argtypes = [C.c_int]
    L.oracle_make_map.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint32, dp, C.c_int]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleFilter:
    """One reference filter instance (EKF or UKF), same call order as the reference's Filter:
    construct(readParams) -> init -> update per step (localization_node.cpp:28-131)."""

    def __init__(self, kind: int, params: OracleParams | None = None, max_landmarks: int = 64, _handle=None):
        self.kind = kind
        self.params = params or make_params()
        self.max_landmarks = max_landmarks
        self.base = 3 if kind in (EKF_SLAM, NAIVE) else 4
        self._h = _handle if _handle is not None else lib().oracle_create(kind, C.byref(self.params), max_landmarks)
        if not self._h:
            raise RuntimeError("Invalid filter choice")

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def init(self, x0: float, y0: float, yaw0: float):
        lib().oracle_init(self._h, x0, y0, yaw0)

    def set_map(self, lm_xy):
        """UKF_LOC: the true map (filter.h:68), sent as the /truth/landmarks wire format float32 [id, x, y]*."""
        lm = np.asarray(lm_xy, dtype=np.float64).reshape(-1, 2)
        m = np.zeros((len(lm), 3), dtype=np.float32)
        m[:, 0] = np.arange(len(lm)); m[:, 1:] = lm.astype(np.float32)
        lib().oracle_set_map(self._h, _fp(np.ascontiguousarray(m.reshape(-1))), len(lm))

    def update(self, fwd: float, ang: float, meas, mode: int = DENSE) -> int:
        m = np.ascontiguousarray(np.asarray(meas, dtype=np.float32).reshape(-1))
        return lib().oracle_update(self._h, np.float32(fwd), np.float32(ang), _fp(m), m.size // 3, mode)

    def predict(self, fwd: float, ang: float, mode: int = DENSE) -> int:
        return lib().oracle_predict(self._h, np.float32(fwd), np.float32(ang), mode)

    def measure(self, meas, mode: int = DENSE) -> int:
        m = np.ascontiguousarray(np.asarray(meas, dtype=np.float32).reshape(-1))
        return lib().oracle_measure(self._h, _fp(m), m.size // 3, mode)

    def set_state(self, x, P, ids, timestep: int = 0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lib().oracle_set_state(self._h, _dp(x), _dp(P), _ip(ids), ids.size, timestep)

    @property
    def n(self) -> int:
        return lib().oracle_state_dim(self._h)

    @property
    def M(self) -> int:
        return lib().oracle_num_landmarks(self._h)

    @property
    def timestep(self) -> int:
        return lib().oracle_timestep(self._h)

    @property
    def status(self) -> int:
        return lib().oracle_status(self._h)

    def state(self) -> np.ndarray:
        x = np.zeros(self.n)
        lib().oracle_get_state(self._h, _dp(x))
        return x

    def cov(self) -> np.ndarray:
        n = self.n
        P = np.zeros((n, n))
        lib().oracle_get_cov(self._h, _dp(P))
        return P

    def landmark_ids(self) -> np.ndarray:
        ids = np.zeros(max(self.M, 1), dtype=np.int32)
        lib().oracle_get_landmark_ids(self._h, _ip(ids))
        return ids[: self.M]

    def assoc_log(self) -> np.ndarray:
        buf = np.zeros(4096, dtype=np.int32)
        k = lib().oracle_get_assoc_log(self._h, _ip(buf), buf.size)
        return buf[:k].copy()

    def sigma_points(self) -> np.ndarray:
        """UKF sigma points X, shape (2n+1, n): row j = sigma point j (ukf.cpp:91-99 wire order)."""
        n = lib().oracle_sigma_rows(self._h)   # X keeps the row count of the last step's start (ukf.cpp:167-171)
        buf = np.zeros((2 * n + 1) * max(n, 1))
        if n:
            lib().oracle_get_sigma_points(self._h, _dp(buf))
        return buf.reshape(2 * n + 1, n)


def philox(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(c0, c1, c2, c3, k0, k1, out)
    return [int(v) for v in out]


def uniform(hi, lo) -> float:
    return lib().oracle_uniform(hi, lo)


def sim_step(params: OracleParams, truth: np.ndarray, fwd, ang, lm_xy: np.ndarray, seed: int, instance: int, step: int):
    """sim_node.py:209-250.  truth (3,) float64 advanced in place; returns float32 [k,3]."""
    lm = np.ascontiguousarray(lm_xy, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(3 * max(len(lm), 1), dtype=np.float32)
    k = lib().oracle_sim_step(C.byref(params), _dp(truth), np.float32(fwd), np.float32(ang), _dp(lm), len(lm),
                              seed, instance, step, _fp(out), len(lm))
    return out[: 3 * k].reshape(k, 3).copy()


def tsp_trajectory(params: OracleParams, lm_xy, landmark_noise, visitation_threshold, bound, pose0, T, seed, instance):
    """sim_node.py:63-152 for one Monte-Carlo instance; returns (fwd float32[T], ang float32[T])."""
    lm = np.ascontiguousarray(lm_xy, dtype=np.float64).reshape(-1, 2)
    fwd = np.zeros(T, dtype=np.float32)
    ang = np.zeros(T, dtype=np.float32)
    rc = lib().oracle_tsp_trajectory(C.byref(params), _dp(lm), len(lm), landmark_noise, visitation_threshold, bound,
                                     pose0[0], pose0[1], pose0[2], T, seed, instance, _fp(fwd), _fp(ang))
    if rc != 0:
        raise ValueError("oracle_tsp_trajectory: bad arguments")
    return fwd, ang


def make_map(map_type: str, n_landmarks: int, bound: float, grid_step: float, min_sep: float, seed: int, instance: int, cap: int = 4096):
    """sim_node.py:155-206 ("grid" | "random"); returns lm_xy [n,2]."""
    code = {"grid": 0, "random": 1, "rand": 1}.get(map_type, -1)
    buf = np.zeros((cap, 2))
    n = lib().oracle_make_map(code, n_landmarks, bound, grid_step, min_sep, seed, instance, _dp(buf), cap)
    if n < 0:
        raise ValueError("Invalid map_type provided." if n == -1 else "map could not be completed")
    return buf[:n].copy()


def run_instance(kind, params, lm_xy, cmd_fwd, cmd_ang, seed, instance, max_landmarks, mode=STRUCTURED, keep=False):
    lm = np.ascontiguousarray(lm_xy, dtype=np.float64).reshape(-1, 2)
    fwd = np.ascontiguousarray(cmd_fwd, dtype=np.float32)
    ang = np.ascontiguousarray(cmd_ang, dtype=np.float32)
    T = len(fwd)
    pose = np.zeros((T, 3))
    truth = np.zeros((T, 3))
    h = C.c_void_p()
    st = lib().oracle_run_instance(kind, C.byref(params), _dp(lm), len(lm), _fp(fwd), _fp(ang), T, seed, instance,
                                   max_landmarks, mode, _dp(pose), _dp(truth), C.byref(h) if keep else None)
    filt = OracleFilter(kind, params, max_landmarks, _handle=h.value) if keep else None
    return st, pose, truth, filt


def bench(kind, params, lm_xy, cmd_fwd, cmd_ang, seed, n_threads, per_thread, max_landmarks, mode=DENSE):
    lm = np.ascontiguousarray(lm_xy, dtype=np.float64).reshape(-1, 2)
    fwd = np.ascontiguousarray(cmd_fwd, dtype=np.float32)
    ang = np.ascontiguousarray(cmd_ang, dtype=np.float32)
    upd = C.c_longlong()
    secs = lib().oracle_bench(kind, C.byref(params), _dp(lm), len(lm), _fp(fwd), _fp(ang), len(fwd), seed,
                              n_threads, per_thread, max_landmarks, mode, C.byref(upd))
    return secs, upd.value


def eigh(A: np.ndarray):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    w = np.zeros(n)
    V = np.zeros((n, n))
    lib().oracle_eigh(_dp(A), n, _dp(w), _dp(V))
    return w, V


def set_trig_mode(mode: int):
    lib().oracle_set_trig_mode(mode)
