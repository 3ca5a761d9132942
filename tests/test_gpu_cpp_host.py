"""The C++ host mirror (live_ekf_slam_b200/host/filter.hpp) driven by the headless iterate() equivalent, compiled
with g++ against the C-ABI library and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "localization_headless")
    pkg = os.path.join(ROOT, "live_ekf_slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(pkg, "host", "localization_headless.cpp"), "-o", exe,
                           "-L" + pkg, "-lslam_filter", "-Wl,-rpath," + pkg])
    return exe


@pytest.mark.parametrize("choice", ["ekf_slam", "ukf_slam"])
def test_cpp_filter_classes(tmp_path, oracle, choice):
    exe = _build(tmp_path)
    p, lm, fwd, ang = H.config2(seed=4, steps=80, filt=choice)
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=2, instance=0)
    lines = [f"{choice} {len(fwd)}"]
    for t in range(len(fwd)):
        m = stream[t]
        lines.append(" ".join([repr(float(fwd[t])), repr(float(ang[t])), str(len(m))] + [repr(float(v)) for v in m.reshape(-1)]))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    okind = oracle.EKF_SLAM if choice == "ekf_slam" else oracle.UKF_SLAM
    of = oracle.OracleFilter(okind, op, 50)
    of.init(0, 0, 0)
    for t in range(len(fwd)):
        of.update(fwd[t], ang[t], stream[t])
        ts, x, y, yaw, M = out[t].split()
        xo = of.state()
        yaw_o = xo[2] if choice == "ekf_slam" else np.arctan2(xo[3], xo[2])
        assert int(ts) == t + 1 and int(M) == of.M
        assert abs(float(x) - xo[0]) <= H.REL_TOL and abs(float(y) - xo[1]) <= H.REL_TOL and abs(float(yaw) - yaw_o) <= H.REL_TOL
    last = out[len(fwd)].split()
    assert abs(float(last[1]) - np.trace(of.cov())) <= 1e-9 * max(1.0, np.trace(of.cov()))
    assert [int(v) for v in last[3:]] == list(of.landmark_ids())


def test_cpp_invalid_choice(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], input="particle 0\n", capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid filter choice" in r.stderr
