"""The C-ABI library: it loads without a GPU, exports every symbol include/slam_filter.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from live_ekf_slam_b200 import shim
from live_ekf_slam_b200.params import Params, SlamParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slam_filter.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slam_[a-z_0-9]+)\s*\(", text)))


def test_header_and_shim_agree():
    assert _declared_symbols() == sorted(shim.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(shim.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = shim.load()
    for name in _declared_symbols():
        assert hasattr(L, name), name


def test_params_struct_layout_matches_header():
    # float x4, double x4, int, float, int, (pad), double x5  -> 104 bytes with natural alignment
    assert C.sizeof(SlamParams) == 104
    assert SlamParams.V_00.offset == 16 and SlamParams.landmark_id_is_known.offset == 48
    assert SlamParams.d_max.offset == 64 and SlamParams.fov_max.offset == 96
    c = Params().to_c()
    assert abs(c.W_11 - 0.01) < 1e-15 and c.compat_noise_bug == 1 and c.landmark_id_is_known == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(shim.SlamError) as e:
        shim.FilterBatch(shim.EKF_SLAM, Params().to_c(), 1, 4, 2)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_invalid_filter_choice_message():
    L = shim.load()
    h = C.c_void_p()
    rc = L.slam_create(7, C.byref(Params().to_c()), 1, 4, 2, 0, C.byref(h))
    assert rc != 0 and b"Invalid filter choice" in L.slam_last_error(None)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError):
        shim.load(str(tmp_path / "nope.so"))
