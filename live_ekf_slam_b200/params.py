"""params.yaml schema of the reference (BP/config/params.yaml:1-113), restricted to the keys the
filter hot path and the measurement generator read.

`load_params` accepts the reference's own yaml file unchanged; `Params.filter` is the reference's
`filter:` switch (localization_node.cpp:33-45).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, asdict

FILTER_CHOICES = ("ekf_slam", "ukf_slam", "ukf_loc", "pose_graph")  # params.yaml:10


class SlamParams(C.Structure):
    """Mirror of `slam_params` in include/slam_filter.h (field order is the ABI)."""
    _fields_ = [
        ("v_d", C.c_float), ("v_th", C.c_float), ("w_r", C.c_float), ("w_b", C.c_float),
        ("V_00", C.c_double), ("V_11", C.c_double), ("W_00", C.c_double), ("W_11", C.c_double),
        ("landmark_id_is_known", C.c_int), ("min_landmark_separation", C.c_float),
        ("compat_noise_bug", C.c_int),
        ("d_max", C.c_double), ("th_max", C.c_double), ("range_max", C.c_double),
        ("fov_min", C.c_double), ("fov_max", C.c_double),
    ]


@dataclass
class Params:
    filter: str = "ekf_slam"                 # params.yaml:11 (default there is "pose_graph", out of scope)
    dt: float = 0.05                         # :14
    num_iterations: int = 1000               # :15
    init_pose: tuple = (0.0, 0.0, 0.0)       # :19-22
    d_max: float = 0.1                       # :27
    th_max: float = 0.0546                   # :28
    range_max: float = 3.0                   # :30
    fov_min: float = -1.57                   # :31
    fov_max: float = 1.57                    # :32
    landmark_id_is_known: bool = True        # :35
    min_landmark_separation: float = 0.1     # :36
    v_d: float = 0.0                         # :41
    v_th: float = 0.0                        # :42
    V_00: float = 0.01                       # :44
    V_11: float = 0.001                      # :45
    w_r: float = 0.0                         # :48
    w_b: float = 0.0                         # :49
    W_00: float = 0.01                       # :51
    W_11: float = 0.01                       # :52
    map_bound: float = 10.0                  # :70
    map_num_landmarks: int = 20              # :71
    map_min_landmark_separation: float = 0.05  # :72
    map_grid_step: float = 4                 # :73
    landmark_noise: float = 0.2              # :90
    visitation_threshold: float = 3.0        # :91
    compat_noise_bug: bool = True            # reproduce filter.h:116-117 (SURVEY B-1)
    extra: dict = field(default_factory=dict)

    def to_c(self) -> SlamParams:
        return SlamParams(
            v_d=self.v_d, v_th=self.v_th, w_r=self.w_r, w_b=self.w_b,
            V_00=self.V_00, V_11=self.V_11, W_00=self.W_00, W_11=self.W_11,
            landmark_id_is_known=int(self.landmark_id_is_known),
            min_landmark_separation=self.min_landmark_separation,
            compat_noise_bug=int(self.compat_noise_bug),
            d_max=self.d_max, th_max=self.th_max, range_max=self.range_max,
            fov_min=self.fov_min, fov_max=self.fov_max)

    def as_dict(self) -> dict:
        d = asdict(self)
        d.pop("extra")
        return d


def from_yaml_dict(cfg: dict) -> Params:
    """Same key paths the reference reads (filter.h:105-121, sim_node.py:216-241, :82-135, :167-185)."""
    p = Params()
    p.filter = str(cfg.get("filter", p.filter))
    if p.filter not in FILTER_CHOICES:
        raise RuntimeError("Invalid filter choice in params.yaml.")  # localization_node.cpp:44
    p.dt = float(cfg.get("dt", p.dt))
    p.num_iterations = int(cfg.get("num_iterations", p.num_iterations))
    ip = cfg.get("init_pose", {})
    p.init_pose = (float(ip.get("x", 0.0)), float(ip.get("y", 0.0)), float(ip.get("yaw", 0.0)))
    con = cfg.get("constraints", {})
    cmd, vis, meas = con.get("commands", {}), con.get("vision", {}), con.get("measurements", {})
    p.d_max = float(cmd.get("d_max", p.d_max)); p.th_max = float(cmd.get("th_max", p.th_max))
    p.range_max = float(vis.get("range_max", p.range_max))
    p.fov_min = float(vis.get("fov_min", p.fov_min)); p.fov_max = float(vis.get("fov_max", p.fov_max))
    p.landmark_id_is_known = bool(meas.get("landmark_id_is_known", p.landmark_id_is_known))
    p.min_landmark_separation = float(meas.get("min_landmark_separation", p.min_landmark_separation))
    pn, sn = cfg.get("process_noise", {}), cfg.get("sensing_noise", {})
    p.v_d = float(pn.get("mean", {}).get("v_d", 0.0)); p.v_th = float(pn.get("mean", {}).get("v_th", 0.0))
    p.V_00 = float(pn.get("cov", {}).get("V_00", p.V_00)); p.V_11 = float(pn.get("cov", {}).get("V_11", p.V_11))
    p.w_r = float(sn.get("mean", {}).get("w_r", 0.0)); p.w_b = float(sn.get("mean", {}).get("w_b", 0.0))
    p.W_00 = float(sn.get("cov", {}).get("W_00", p.W_00)); p.W_11 = float(sn.get("cov", {}).get("W_11", p.W_11))
    m = cfg.get("map", {})
    p.map_bound = float(m.get("bound", p.map_bound))
    p.map_num_landmarks = int(m.get("num_landmarks", p.map_num_landmarks))
    p.map_min_landmark_separation = float(m.get("min_landmark_separation", p.map_min_landmark_separation))
    p.map_grid_step = float(m.get("grid_step", p.map_grid_step))
    tg = cfg.get("trajectory_gen", {})
    p.landmark_noise = float(tg.get("landmark_noise", p.landmark_noise))
    p.visitation_threshold = float(tg.get("visitation_threshold", p.visitation_threshold))
    p.compat_noise_bug = bool(cfg.get("compat_noise_bug", True))
    return p


def load_params(path: str) -> Params:
    import yaml
    with open(path) as f:
        return from_yaml_dict(yaml.safe_load(f))
