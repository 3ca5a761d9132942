"""ctypes binding of the C-ABI declared in include/slam_filter.h (libslam_filter.so, built in-tree by
`live_ekf_slam_b200/csrc/Makefile`).  This is the stub a maintainer of the reference's Python/ROS nodes
would add.  There is no CPU fallback: if the CUDA library is missing or no B200 is
visible, loading / slam_create fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .params import SlamParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libslam_filter.so")

EKF_SLAM = 1   # FilterChoice::EKF_SLAM, filter.h:46
UKF_LOC = 2    # FilterChoice::UKF_LOC, filter.h:47 (localisation on the true map)
UKF_SLAM = 3   # FilterChoice::UKF_SLAM, filter.h:48
NAIVE = 5      # FilterChoice::NAIVE_COMMAND_PROPAGATION, filter.h:50 (NaiveFilter, filter.h:325-370)
STATUS_NAN, STATUS_SAME_STEP_REMATCH, STATUS_CAPACITY, STATUS_MEAS_OVERFLOW, STATUS_BAD_ID = 1, 2, 4, 8, 16
NUM_STATS = 14

#: every symbol include/slam_filter.h declares (tests check the built library exports all of them)
ABI_SYMBOLS = (
    "slam_create", "slam_destroy", "slam_set_map", "slam_last_error", "slam_stream", "slam_synchronize", "slam_batch", "slam_kind",
    "slam_init", "slam_step", "slam_step_device", "slam_predict", "slam_update", "slam_predict_device",
    "slam_update_device", "slam_get_timestep", "slam_get_num_landmarks", "slam_get_status", "slam_get_state",
    "slam_get_state_vector", "slam_get_cov", "slam_get_landmark_ids", "slam_get_assoc", "slam_get_sigma_points",
    "slam_get_poses", "slam_get_all_status", "slam_get_all_num_landmarks", "slam_set_state",
    "slam_sim_create", "slam_sim_destroy", "slam_sim_make_trajectories", "slam_sim_make_maps", "slam_sim_get_map", "slam_sim_reset", "slam_sim_step", "slam_sim_step_device",
    "slam_sim_meas", "slam_sim_n_meas", "slam_sim_get_truth", "slam_sim_get_meas",
    "slam_run", "slam_run_device", "slam_reset", "slam_step_io", "slam_run_io", "slam_set_profiling", "slam_get_profile",
    "slam_accumulate_error", "slam_get_stats", "slam_reset_stats", "slam_get_error_histogram",
    "slam_kernel_launches", "slam_build_info", "slam_tune", "slam_get_ukf_routes",
)

_lib = None


def load(path: str | None = None):
    """dlopen the CUDA library and declare argument types.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `make -C live_ekf_slam_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.slam_create.argtypes = [C.c_int, C.POINTER(SlamParams), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.slam_destroy.argtypes = [vp]
    L.slam_set_map.argtypes = [vp, fp, C.c_int]
    L.slam_last_error.argtypes = [vp]
    L.slam_last_error.restype = C.c_char_p
    L.slam_stream.argtypes = [vp]
    L.slam_stream.restype = vp
    L.slam_synchronize.argtypes = [vp]
    L.slam_batch.argtypes = [vp]
    L.slam_kind.argtypes = [vp]
    L.slam_init.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    L.slam_step.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.slam_step_device.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.slam_predict.argtypes = [vp, vp, vp, C.c_int]
    L.slam_update.argtypes = [vp, vp, vp]
    L.slam_predict_device.argtypes = [vp, vp, vp, C.c_int]
    L.slam_update_device.argtypes = [vp, vp, vp]
    for name in ("slam_get_timestep", "slam_get_num_landmarks", "slam_get_status"):
        getattr(L, name).argtypes = [vp, C.c_int, ip]
    for name in ("slam_get_state", "slam_get_state_vector", "slam_get_cov", "slam_get_sigma_points"):
        getattr(L, name).argtypes = [vp, C.c_int, dp, ip]
    L.slam_get_landmark_ids.argtypes = [vp, C.c_int, ip, ip]
    L.slam_get_assoc.argtypes = [vp, C.c_int, ip, ip]
    L.slam_get_poses.argtypes = [vp, dp]
    L.slam_get_all_status.argtypes = [vp, ip]
    L.slam_get_all_num_landmarks.argtypes = [vp, ip]
    L.slam_set_state.argtypes = [vp, C.c_int, dp, dp, ip, C.c_int, C.c_int]
    L.slam_sim_create.argtypes = [vp, dp, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(vp)]
    L.slam_sim_destroy.argtypes = [vp]
    L.slam_sim_make_trajectories.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, vp, vp]
    L.slam_sim_reset.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.slam_sim_make_maps.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, ip]
    L.slam_sim_get_map.argtypes = [vp, C.c_int, dp, ip]
    L.slam_sim_step.argtypes = [vp, vp, vp, C.c_int, C.c_uint32]
    L.slam_sim_step_device.argtypes = [vp, vp, vp, C.c_int, C.c_uint32]
    L.slam_sim_meas.argtypes = [vp]
    L.slam_sim_meas.restype = vp
    L.slam_sim_n_meas.argtypes = [vp]
    L.slam_sim_n_meas.restype = vp
    L.slam_sim_get_truth.argtypes = [vp, dp]
    L.slam_sim_get_meas.argtypes = [vp, fp, ip]
    L.slam_run.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_uint32]
    L.slam_run_device.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_uint32]
    L.slam_reset.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    L.slam_step_io.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
    L.slam_run_io.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int]
    L.slam_set_profiling.argtypes = [vp, C.c_int]
    L.slam_get_profile.argtypes = [vp, dp, C.POINTER(C.c_longlong)]
    L.slam_accumulate_error.argtypes = [vp, vp]
    L.slam_get_stats.argtypes = [vp, dp]
    L.slam_reset_stats.argtypes = [vp]
    L.slam_get_error_histogram.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_longlong), dp]
    L.slam_kernel_launches.argtypes = [vp]
    L.slam_kernel_launches.restype = C.c_longlong
    L.slam_build_info.argtypes = [C.c_char_p, C.c_int]
    L.slam_tune.argtypes = [vp, C.c_int, C.c_int]
    L.slam_get_ukf_routes.argtypes = [vp, C.POINTER(C.c_longlong)]
    if path is None:
        _lib = L
    return L


def _ptr(a):
    """host numpy array / torch tensor / raw integer address -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def pack_meas(batch: int, max_meas: int, per_instance) -> tuple[np.ndarray, np.ndarray]:
    """list (one per instance) of [k,3] float32 [id, r, b] arrays -> the ABI's [batch][max_meas][3] buffer + counts.
    Counts are NOT clamped: a count above max_meas makes the kernels flag SLAM_STATUS_MEAS_OVERFLOW."""
    if len(per_instance) != batch:
        raise ValueError("one measurement list per instance expected")
    meas = np.zeros((batch, max_meas, 3), dtype=np.float32)
    n = np.zeros(batch, dtype=np.int32)
    for i, m in enumerate(per_instance):
        m = np.asarray(m, dtype=np.float32).reshape(-1, 3)
        k = min(len(m), max_meas)
        meas[i, :k] = m[:k]
        n[i] = len(m)
    return meas, n


class SlamError(RuntimeError):
    """The C++ reference throws std::runtime_error (filter.h:5, localization_node.cpp:44); the ABI returns a
    status and the shim re-raises."""


class FilterBatch:
    """A batch of independent filter instances on one GPU (thin wrapper over a slam_handle_t)."""

    def __init__(self, kind: int, params: SlamParams, batch: int = 1, max_landmarks: int = 50, max_meas: int = 8,
                 device: int = 0):
        self._L = load()
        self._h = C.c_void_p()
        self.kind, self.batch, self.max_landmarks, self.max_meas, self.device = kind, batch, max_landmarks, max_meas, device
        self.base = 3 if kind in (EKF_SLAM, NAIVE) else 4
        self._params = params
        rc = self._L.slam_create(kind, C.byref(params), batch, max_landmarks, max_meas, device, C.byref(self._h))
        if rc != 0:
            msg = self._L.slam_last_error(None).decode()
            self._h = C.c_void_p()
            raise SlamError(msg)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.slam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SlamError(self._L.slam_last_error(self._h).decode())

    @property
    def stream(self) -> int:
        return int(self._L.slam_stream(self._h) or 0)

    def synchronize(self):
        self._ck(self._L.slam_synchronize(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._L.slam_kernel_launches(self._h))

    def set_map(self, lm_xy):
        """UKF_LOC: the true map the reference receives on /truth/landmarks and keeps in Filter::map (filter.h:68):
        float32 [id, x, y]* with id == index.  Accepts [N,2] coordinates or the flat wire format."""
        a = np.asarray(lm_xy)
        if a.ndim == 2 and a.shape[1] == 2:
            m = np.zeros((len(a), 3), dtype=np.float32)
            m[:, 0] = np.arange(len(a)); m[:, 1:] = a.astype(np.float32)
        else:
            m = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
        m = np.ascontiguousarray(m.reshape(-1))
        self._ck(self._L.slam_set_map(self._h, m.ctypes.data_as(C.POINTER(C.c_float)), m.size // 3))

    # -- Filter interface (filter.h:59-61)
    def init(self, x_0: float, y_0: float, yaw_0: float):
        self._ck(self._L.slam_init(self._h, x_0, y_0, yaw_0))

    def _cmd(self, fwd, ang):
        f = np.ascontiguousarray(np.atleast_1d(np.asarray(fwd, dtype=np.float32)))
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(ang, dtype=np.float32)))
        stride = 0 if f.size == 1 else 1
        if stride and f.size != self.batch:
            raise ValueError("fwd/ang must be scalars or have one entry per instance")
        return f, a, stride

    def pack_meas(self, per_instance) -> tuple[np.ndarray, np.ndarray]:
        """list (one per instance) of [k,3] float32 arrays -> the [batch][max_meas][3] buffer + counts."""
        return pack_meas(self.batch, self.max_meas, per_instance)

    def step(self, fwd, ang, meas: np.ndarray, n_meas: np.ndarray):
        """Filter::update for every instance, host buffers."""
        f, a, stride = self._cmd(fwd, ang)
        meas = np.ascontiguousarray(meas, dtype=np.float32)
        n_meas = np.ascontiguousarray(n_meas, dtype=np.int32)
        assert meas.size == self.batch * self.max_meas * 3 and n_meas.size == self.batch
        self._ck(self._L.slam_step(self._h, _ptr(f), _ptr(a), stride, _ptr(meas), _ptr(n_meas)))

    def step_device(self, d_fwd, d_ang, cmd_stride: int, d_meas, d_n_meas):
        self._ck(self._L.slam_step_device(self._h, _ptr(d_fwd), _ptr(d_ang), cmd_stride, _ptr(d_meas), _ptr(d_n_meas)))

    def predict(self, fwd, ang):
        f, a, stride = self._cmd(fwd, ang)
        self._ck(self._L.slam_predict(self._h, _ptr(f), _ptr(a), stride))

    def update(self, meas: np.ndarray, n_meas: np.ndarray):
        meas = np.ascontiguousarray(meas, dtype=np.float32)
        n_meas = np.ascontiguousarray(n_meas, dtype=np.int32)
        self._ck(self._L.slam_update(self._h, _ptr(meas), _ptr(n_meas)))

    # -- getters
    def _int(self, fn, inst):
        v = C.c_int()
        self._ck(fn(self._h, inst, C.byref(v)))
        return v.value

    def timestep(self, inst=0) -> int:
        return self._int(self._L.slam_get_timestep, inst)

    def num_landmarks(self, inst=0) -> int:
        return self._int(self._L.slam_get_num_landmarks, inst)

    def status(self, inst=0) -> int:
        return self._int(self._L.slam_get_status, inst)

    def state(self, inst=0) -> np.ndarray:
        x = np.zeros(self.base + 2 * self.max_landmarks)
        n = C.c_int()
        self._ck(self._L.slam_get_state(self._h, inst, x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return x[: n.value].copy()

    def state_vector(self, inst=0) -> np.ndarray:
        x = np.zeros(self.base + 2 * self.max_landmarks)
        n = C.c_int()
        self._ck(self._L.slam_get_state_vector(self._h, inst, x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return x[: n.value].copy()

    def cov(self, inst=0) -> np.ndarray:
        nmax = self.base + 2 * self.max_landmarks
        P = np.zeros(nmax * nmax)
        n = C.c_int()
        self._ck(self._L.slam_get_cov(self._h, inst, P.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return P[: n.value * n.value].reshape(n.value, n.value).copy()

    def landmark_ids(self, inst=0) -> np.ndarray:
        ids = np.zeros(self.max_landmarks, dtype=np.int32)
        m = C.c_int()
        self._ck(self._L.slam_get_landmark_ids(self._h, inst, ids.ctypes.data_as(C.POINTER(C.c_int)), C.byref(m)))
        return ids[: m.value].copy()

    def assoc(self, inst=0) -> np.ndarray:
        a = np.zeros(self.max_meas, dtype=np.int32)
        k = C.c_int()
        self._ck(self._L.slam_get_assoc(self._h, inst, a.ctypes.data_as(C.POINTER(C.c_int)), C.byref(k)))
        return a[: k.value].copy()

    def sigma_points(self, inst=0) -> np.ndarray:
        """UKF: X of the last update, shape (2n+1, n): row j = sigma point j, i.e. the flattened array is UKFState.X
        exactly as ukf.cpp:91-99 fills it (n = 4 + 2M of that update's prior; 4 x 9 zeros before the first update)."""
        nmax = self.base + 2 * self.max_landmarks
        X = np.zeros(nmax * (2 * nmax + 1))
        n = C.c_int()
        self._ck(self._L.slam_get_sigma_points(self._h, inst, X.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return X[: n.value * (2 * n.value + 1)].reshape(2 * n.value + 1, n.value).copy()

    def poses(self) -> np.ndarray:
        out = np.zeros((self.batch, 3))
        self._ck(self._L.slam_get_poses(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def all_status(self) -> np.ndarray:
        out = np.zeros(self.batch, dtype=np.int32)
        self._ck(self._L.slam_get_all_status(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def all_num_landmarks(self) -> np.ndarray:
        out = np.zeros(self.batch, dtype=np.int32)
        self._ck(self._L.slam_get_all_num_landmarks(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def set_state(self, inst, x, P, ids, timestep=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        self._ck(self._L.slam_set_state(self._h, inst, x.ctypes.data_as(dp), P.ctypes.data_as(dp),
                                        ids.ctypes.data_as(ip), int(ids.size), timestep))

    def reset(self, x_0=0.0, y_0=0.0, yaw_0=0.0):
        """Filter::init on the device, asynchronous (between Monte-Carlo sweeps)."""
        self._ck(self._L.slam_reset(self._h, x_0, y_0, yaw_0))

    def step_io(self, fwd, ang, cmd_stride: int, meas, n_meas, poses_out):
        """slam_step_io with caller-owned (ideally pinned) host buffers; asynchronous."""
        self._ck(self._L.slam_step_io(self._h, _ptr(fwd), _ptr(ang), cmd_stride, _ptr(meas), _ptr(n_meas), _ptr(poses_out)))

    def run_io(self, fwd, ang, cmd_stride: int, meas, n_meas, poses_out, T: int):
        """slam_run_io: a whole recorded run through HOST buffers (numpy arrays, pinned torch tensors or raw
        addresses); asynchronous -- synchronize() before reading poses_out."""
        self._ck(self._L.slam_run_io(self._h, _ptr(fwd), _ptr(ang), cmd_stride, _ptr(meas), _ptr(n_meas), _ptr(poses_out), T))

    def ukf_routes(self) -> np.ndarray:
        """instance-steps taken by the dense (generation 3), QL (generation 2) and explicit-eigenvector / rescue routes"""
        out = np.zeros(3, dtype=np.int64)
        self._ck(self._L.slam_get_ukf_routes(self._h, out.ctypes.data_as(C.POINTER(C.c_longlong))))
        return out

    def tune(self, key: int, value: int):
        self._ck(self._L.slam_tune(self._h, key, value))

    def set_profiling(self, on: bool):
        self._ck(self._L.slam_set_profiling(self._h, int(on)))

    def profile(self):
        ms = C.c_double()
        n = C.c_longlong()
        self._ck(self._L.slam_get_profile(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- stats
    def stats(self) -> np.ndarray:
        out = np.zeros(NUM_STATS)
        self._ck(self._L.slam_get_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def reset_stats(self):
        self._ck(self._L.slam_reset_stats(self._h))

    def error_histogram(self, lo: float, hi: float, nbins: int, per_instance: bool = False):
        """Histogram of the per-run average position error (plotting_node.py:195-218) over the batch: int64
        counts[nbins + 2] (under / bins / over) and, optionally, the per-instance averages (the ekf.csv column of the
        reference's recorded runs, make_bar_graphs.py:11-18)."""
        counts = np.zeros(nbins + 2, dtype=np.int64)
        avg = np.zeros(self.batch) if per_instance else None
        self._ck(self._L.slam_get_error_histogram(self._h, lo, hi, nbins, counts.ctypes.data_as(C.POINTER(C.c_longlong)),
                                                  avg.ctypes.data_as(C.POINTER(C.c_double)) if avg is not None else None))
        return (counts, avg) if per_instance else counts


class Simulator:
    """On-GPU measurement generator bound to a FilterBatch (sim_node.py:209-250)."""

    def __init__(self, filt: FilterBatch, lm_xy: np.ndarray, seed: int = 0, instance_offset: int = 0):
        self._f = filt
        self._L = filt._L
        self._s = C.c_void_p()
        lm = np.ascontiguousarray(lm_xy, dtype=np.float64).reshape(-1, 2)
        self.n_lm = len(lm)
        filt._ck(self._L.slam_sim_create(filt._h, lm.ctypes.data_as(C.POINTER(C.c_double)), len(lm), seed,
                                         instance_offset, C.byref(self._s)))

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            self._L.slam_sim_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, x_0=0.0, y_0=0.0, yaw_0=0.0):
        self._f._ck(self._L.slam_sim_reset(self._s, x_0, y_0, yaw_0))

    def step(self, fwd, ang, step: int):
        f, a, stride = self._f._cmd(fwd, ang)
        self._f._ck(self._L.slam_sim_step(self._s, _ptr(f), _ptr(a), stride, step))

    def step_device(self, d_fwd, d_ang, cmd_stride: int, step: int):
        self._f._ck(self._L.slam_sim_step_device(self._s, _ptr(d_fwd), _ptr(d_ang), cmd_stride, step))

    @property
    def d_meas(self) -> int:
        return int(self._L.slam_sim_meas(self._s))

    @property
    def d_n_meas(self) -> int:
        return int(self._L.slam_sim_n_meas(self._s))

    def truth(self) -> np.ndarray:
        out = np.zeros((self._f.batch, 3))
        self._f._ck(self._L.slam_sim_get_truth(self._s, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def meas(self):
        m = np.zeros((self._f.batch, self._f.max_meas, 3), dtype=np.float32)
        n = np.zeros(self._f.batch, dtype=np.int32)
        self._f._ck(self._L.slam_sim_get_meas(self._s, m.ctypes.data_as(C.POINTER(C.c_float)),
                                              n.ctypes.data_as(C.POINTER(C.c_int))))
        return m, n

    def run(self, cmd_fwd, cmd_ang, first_step: int = 0):
        """slam_run: T fused sim+filter steps on the device; commands shared ([T]) or per instance ([T][batch])."""
        f = np.ascontiguousarray(cmd_fwd, dtype=np.float32)
        a = np.ascontiguousarray(cmd_ang, dtype=np.float32)
        stride = 0 if f.ndim == 1 else 1
        T = f.shape[0]
        self._f._ck(self._L.slam_run(self._f._h, self._s, _ptr(f), _ptr(a), stride, T, first_step))

    def run_device(self, d_fwd, d_ang, cmd_stride: int, T: int, first_step: int = 0):
        self._f._ck(self._L.slam_run_device(self._f._h, self._s, _ptr(d_fwd), _ptr(d_ang), cmd_stride, T, first_step))

    def make_trajectories(self, landmark_noise: float, visitation_threshold: float, bound: float, pose0, T: int, d_fwd, d_ang):
        """generate_trajectory (sim_node.py:63-152) for every instance on the device: fills the DEVICE buffers d_fwd / d_ang
        ([T][batch] float32, e.g. torch tensors) with per-instance command trajectories (use with run_device(..., cmd_stride=1))."""
        self._f._ck(self._L.slam_sim_make_trajectories(self._s, landmark_noise, visitation_threshold, bound,
                                                       pose0[0], pose0[1], pose0[2], T, _ptr(d_fwd), _ptr(d_ang)))

    def make_maps(self, map_type: str, n_landmarks: int, bound: float, grid_step: float = 4.0, min_sep: float = 0.05) -> int:
        """generate_landmarks (sim_node.py:155-206) on the device, one map per vehicle: "grid" or "random" / "rand"."""
        code = {"grid": 0, "random": 1, "rand": 1}.get(map_type, -1)
        n = C.c_int()
        self._f._ck(self._L.slam_sim_make_maps(self._s, code, n_landmarks, bound, grid_step, min_sep, C.byref(n)))
        self.n_lm = n.value
        return n.value

    def get_map(self, inst: int = 0) -> np.ndarray:
        lm = np.zeros((self.n_lm, 2))
        n = C.c_int()
        self._f._ck(self._L.slam_sim_get_map(self._s, inst, lm.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return lm[: n.value]

    def accumulate_error(self):
        self._f._ck(self._L.slam_accumulate_error(self._f._h, self._s))
