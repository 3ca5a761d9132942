"""Host-side logic that needs no GPU: params.yaml schema, the filter: switch, workload generators, message packing."""
import math

import numpy as np
import pytest
import yaml

from live_ekf_slam_b200 import Params, from_yaml_dict, workload as wl
from live_ekf_slam_b200 import shim
from live_ekf_slam_b200.parallel import shard_range, weak_offset, derive_accuracy

# the keys of the reference's BP/config/params.yaml that the hot path reads, with its default values
REFERENCE_YAML = """
filter: "ekf_slam"
dt: 0.05
num_iterations: 1000
init_pose: {x: 0.0, y: 0.0, yaw: 0.0}
constraints:
  commands: {d_max: 0.1, th_max: 0.0546}
  vision: {range_max: 3.0, fov_min: -1.57, fov_max: 1.57}
  measurements: {landmark_id_is_known: true, min_landmark_separation: 0.1}
process_noise:
  mean: {v_d: 0.0, v_th: 0.0}
  cov: {V_00: 0.01, V_11: 0.001}
sensing_noise:
  mean: {w_r: 0.0, w_b: 0.0}
  cov: {W_00: 0.01, W_11: 0.01}
ukf: {W_0: 0.2}
map: {bound: 10.0, num_landmarks: 20, min_landmark_separation: 0.05, grid_step: 4}
trajectory_gen: {landmark_noise: 0.2, visitation_threshold: 3.0}
"""


def test_yaml_schema_roundtrip():
    p = from_yaml_dict(yaml.safe_load(REFERENCE_YAML))
    d = Params()
    assert p.as_dict() == d.as_dict()
    assert p.filter == "ekf_slam" and p.landmark_id_is_known is True and p.th_max == 0.0546


def test_invalid_filter_raises_like_reference():
    cfg = yaml.safe_load(REFERENCE_YAML)
    cfg["filter"] = "particle"
    with pytest.raises(RuntimeError, match="Invalid filter choice"):
        from_yaml_dict(cfg)


def test_filter_switch_out_of_scope_choices():
    from live_ekf_slam_b200.filter import make_filter
    with pytest.raises(RuntimeError, match="outside the B200 hot path"):
        make_filter(Params(filter="pose_graph"))


def test_reference_grid_map_is_25_landmarks():
    lm = wl.grid_map(Params())                      # sim_node.py:167-176 with bound 10, grid_step 4
    assert lm.shape == (25, 2)
    assert lm[0].tolist() == [-8.0, -8.0] and lm[1].tolist() == [-8.0, -4.0] and lm[-1].tolist() == [8.0, 8.0]


def test_grid_5x10_and_random_map():
    lm = wl.grid_map_5x10()
    assert lm.shape == (50, 2) and lm[:, 0].min() == -8 and lm[:, 1].max() == 9
    rng = np.random.default_rng(0)
    r = wl.random_map(20, 10.0, 0.05, rng)
    assert r.shape == (20, 2) and np.abs(r).max() <= 10
    dmin = min(math.hypot(*(r[i] - r[j])) for i in range(20) for j in range(i))
    assert dmin >= 0.05
    rf = wl.random_map_fast(300, 10.0, 0.3, np.random.default_rng(1))
    d = np.sqrt(((rf[:, None, :] - rf[None, :, :]) ** 2).sum(-1)) + np.eye(300) * 9
    assert d.min() >= 0.3


def test_tsp_trajectory_respects_command_constraints():
    p = Params()
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, np.random.default_rng(0), 500)
    assert fwd.dtype == np.float32 and ang.dtype == np.float32 and len(fwd) == 500
    assert (fwd >= 0).all() and (fwd <= np.float32(p.d_max)).all() and (np.abs(ang) <= np.float32(p.th_max)).all()
    # deterministic given the seed
    f2, a2 = wl.tsp_trajectory(lm, p, np.random.default_rng(0), 500)
    assert (fwd == f2).all() and (ang == a2).all()


def test_pack_meas_layout_and_ragged_inputs():
    per = [np.array([[3, 1.0, 0.5], [7, 2.0, -0.5]], dtype=np.float32), [], np.zeros((5, 3), dtype=np.float32)]
    meas, n = shim.pack_meas(3, 4, per)
    assert meas.shape == (3, 4, 3) and n.tolist() == [2, 0, 5]      # counts are not clamped (overflow is flagged on device)
    assert meas[0, 1].tolist() == [7.0, 2.0, -0.5] and (meas[1] == 0).all()
    with pytest.raises(ValueError):
        shim.pack_meas(2, 4, per)


def test_sharding_helpers():
    tot = 65536
    for world in (1, 2, 4, 8, 3):
        seen = []
        for r in range(world):
            first, cnt = shard_range(tot, r, world)
            seen.extend(range(first, first + cnt))
            for i in (first, first + cnt - 1):
                assert i * world // tot == r          # instance i on GPU floor(i*G/65536), SURVEY 8d config 5
        assert seen == list(range(tot))
    assert weak_offset(4096, 3) == 12288
    acc = derive_accuracy(np.array([4.0, 4.0, 16.0, 1.0, 2.0, 12.0, 1.0, 10.0, 0, 0, 0, 0]), 2)
    assert acc["rmse_x"] == 1.0 and acc["rmse_y"] == 2.0 and acc["mean_pos_err_m"] == 0.5 and acc["mean_final_landmarks"] == 5.0
