"""CPU-only: the Python params loader and the C++ YAML-subset reader agree on the test fixture and on the reference's key
layout (the C++ side is compiled into a tiny probe; no CUDA call is made)."""
import os
import subprocess

import yaml

from live_ekf_slam_b200.params import from_yaml_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(ROOT, "tests", "golden", "params_node.yaml")

PROBE = r'''
#include <cstdio>
#include "params_yaml.hpp"
int main(int argc, char** argv) {
    try {
        const slam_host::YamlNode c = slam_host::load_yaml_subset(argv[1]);
        const slam_params p = slam_host::read_common_params(c);
        std::printf("%s %.9g %.9g %.9g %.9g %.9g %.17g %.17g %.17g %.17g %d %.9g %d %.17g %.17g %.17g %.17g %.17g %.9g %.9g\n",
                    c["filter"].as_string().c_str(), c["dt"].as_float(), p.v_d, p.v_th, p.w_r, p.w_b, p.V_00, p.V_11, p.W_00, p.W_11,
                    p.landmark_id_is_known, p.min_landmark_separation, p.compat_noise_bug, p.d_max, p.th_max, p.range_max, p.fov_min,
                    p.fov_max, c["init_pose"]["x"].as_float(), c["init_pose"]["yaw"].as_float());
    } catch (const std::runtime_error& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
'''


def test_cpp_yaml_subset_reader_matches_python_loader(tmp_path):
    src = tmp_path / "probe.cpp"
    src.write_text(PROBE)
    exe = str(tmp_path / "probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "live_ekf_slam_b200", "host"), str(src), "-o", exe])
    out = subprocess.run([exe, YAML], capture_output=True, text=True, check=True).stdout.split()
    with open(YAML) as f:
        p = from_yaml_dict(yaml.safe_load(f))
    assert out[0] == p.filter == "ekf_slam"
    got = [float(v) for v in out[1:]]
    want = [p.dt, p.v_d, p.v_th, p.w_r, p.w_b, p.V_00, p.V_11, p.W_00, p.W_11, int(p.landmark_id_is_known), p.min_landmark_separation,
            1, p.d_max, p.th_max, p.range_max, p.fov_min, p.fov_max, p.init_pose[0], p.init_pose[2]]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert abs(g - w) <= 1e-6 * max(1.0, abs(w)), (g, w)
    r = subprocess.run([exe, str(tmp_path / "nope.yaml")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
    partial = tmp_path / "partial.yaml"
    partial.write_text("filter: ekf_slam\ndt: 0.05\n")
    r = subprocess.run([exe, str(partial)], capture_output=True, text=True)
    assert r.returncode == 1 and "missing key" in r.stderr
