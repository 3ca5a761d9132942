"""The C++ host mirror (live_ekf_slam_b200/host/filter.hpp) driven by the headless localization_node equivalent
(host/localization_headless.cpp: params.yaml -> `filter:` switch -> init -> iterate() with publishState()), compiled with
g++ against the C-ABI library and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(ROOT, "tests", "golden", "params_node.yaml")


def _build(tmp_path):
    exe = str(tmp_path / "localization_headless")
    pkg = os.path.join(ROOT, "live_ekf_slam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(pkg, "host", "localization_headless.cpp"), "-o", exe,
                           "-L" + pkg, "-lslam_filter", "-Wl,-rpath," + pkg])
    return exe


def _step_lines(fwd, ang, stream):
    lines = []
    for t in range(len(fwd)):
        m = stream[t]
        lines.append(" ".join([repr(float(fwd[t])), repr(float(ang[t])), str(len(m))] + [repr(float(v)) for v in m.reshape(-1)]))
    return lines


@pytest.mark.parametrize("choice", ["ekf_slam", "ukf_slam"])
def test_cpp_filter_classes(tmp_path, oracle, choice):
    """legacy form: filter choice on stdin, default parameters; the per-tick lines come from the PUBLISHED float32 message"""
    exe = _build(tmp_path)
    p, lm, fwd, ang = H.config2(seed=4, steps=80, filt=choice)
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=2, instance=0)
    lines = [f"{choice} {len(fwd)}"] + _step_lines(fwd, ang, stream)
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
        np.testing.assert_allclose([float(x), float(y), float(yaw)], np.float32([xo[0], xo[1], yaw_o]), rtol=3e-7, atol=1e-9)
    last = out[len(fwd)].split()
    assert abs(float(last[1]) - np.trace(of.cov())) <= 1e-9 * max(1.0, np.trace(of.cov()))
    assert [int(v) for v in last[3:]] == list(of.landmark_ids())
    if choice == "ukf_slam":
        Xo = of.sigma_points().astype(np.float32)
        tag, cnt, sm = out[len(fwd) + 1].split()
        assert tag == "X" and int(cnt) == Xo.size
        assert abs(float(sm) - float(Xo.astype(np.float64).sum())) <= 1e-4 * max(1.0, abs(float(Xo.sum())))


@pytest.mark.parametrize("choice", ["ekf_slam", "ukf_slam", "ukf_loc"])
def test_cpp_headless_node_reads_params_yaml(tmp_path, oracle, choice):
    """yaml-driven form (localization_node.cpp:28-47): the `filter:` key selects the class, readParams takes the noise
    profile, the start pose comes from init_pose; every value of the fixture differs from the reference defaults."""
    import yaml
    exe = _build(tmp_path)
    cfg_path = str(tmp_path / "params.yaml")
    with open(YAML) as f:
        text = f.read().replace('filter: "ekf_slam"', f'filter: "{choice}"')
    with open(cfg_path, "w") as f:
        f.write(text)
    from live_ekf_slam_b200.params import from_yaml_dict
    p = from_yaml_dict(yaml.safe_load(text))
    assert p.filter == choice and p.v_d == 0.002 and p.W_11 == 0.008 and p.init_pose == (0.25, -0.5, 0.1)
    rng = np.random.default_rng(3)
    from live_ekf_slam_b200 import workload as wl
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, rng, 90)
    op = H.oracle_params(oracle, p)
    # the simulated vehicle starts at the yaml's init pose too
    truth = np.array(p.init_pose, dtype=float)
    stream = [oracle.sim_step(op, truth, fwd[t], ang[t], lm, 5, 0, t) for t in range(len(fwd))]
    head = []
    if choice == "ukf_loc":
        m = np.zeros((len(lm), 3), dtype=np.float32)
        m[:, 0] = np.arange(len(lm)); m[:, 1:] = lm.astype(np.float32)
        head.append("map %d " % len(lm) + " ".join(repr(float(v)) for v in m.reshape(-1)))
    lines = head + [str(len(fwd))] + _step_lines(fwd, ang, stream)
    r = subprocess.run([exe, cfg_path, "default"], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    out = r.stdout.splitlines()
    assert f"filter {choice} dt 0.02" in r.stderr
    okind = {"ekf_slam": oracle.EKF_SLAM, "ukf_slam": oracle.UKF_SLAM, "ukf_loc": oracle.UKF_LOC}[choice]
    of = oracle.OracleFilter(okind, op, 50)
    of.init(*p.init_pose)
    if choice == "ukf_loc":
        of.set_map(lm)
    for t in range(len(fwd)):
        of.update(fwd[t], ang[t], stream[t])
        ts, x, y, yaw, M = out[t].split()
        xo = of.state()
        yaw_o = xo[2] if choice == "ekf_slam" else np.arctan2(xo[3], xo[2])
        assert int(ts) == t + 1 and int(M) == of.M
        np.testing.assert_allclose([float(x), float(y), float(yaw)], np.float32([xo[0], xo[1], yaw_o]), rtol=3e-7, atol=1e-9)
    last = out[len(fwd)].split()
    assert abs(float(last[1]) - np.trace(of.cov())) <= 1e-9 * max(1.0, np.trace(of.cov()))
    assert [int(v) for v in last[3:]] == list(of.landmark_ids())


def test_cpp_invalid_choice(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], input="particle 0\n", capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid filter choice" in r.stderr
    bad = tmp_path / "bad.yaml"
    bad.write_text('filter: "pose_graph"\n')
    r = subprocess.run([exe, str(bad)], input="0\n", capture_output=True, text=True)
    assert r.returncode == 1 and "outside the B200 hot path" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "missing.yaml")], input="0\n", capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
