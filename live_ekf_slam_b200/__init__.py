"""B200-native EKF-SLAM / UKF-SLAM filter hot path behind the Filter plugin interface of
kevin-robb/live_ekf_slam (ekf_ws/src/localization_pkg/include/localization_pkg/filter.h:54-145).

The compute lives in hand-written sm_100a CUDA kernels (csrc/) behind the C-ABI declared in
include/slam_filter.h; this package is the ctypes shim and the host-side mirror of the
reference's Filter/EKF/UKF classes.  There is no CPU fallback.
"""
from .params import Params, load_params, from_yaml_dict  # noqa: F401

__version__ = "0.1.0"
