"""Host-side mirror of the reference's Filter plugin interface
(ekf_ws/src/localization_pkg/include/localization_pkg/filter.h:54-223) over the CUDA C-ABI.

Same method names, argument meaning and error behaviour as the C++ classes: construct, `readParams(config)`,
`init(x_0, y_0, yaw_0)`, `update(cmdMsg, lmMeasMsg)` once per timer tick, `publishState()` /
`getStateVector()`.  ROS message types are replaced by plain data: a command is `(fwd, ang)` float32
(Command.msg:3-5) and a measurement message is the flat float32 list `[id, r, b, id, r, b, ...]`
(std_msgs/Float32MultiArray as filled by sim_node.py:245-250).

`make_filter(config)` is the factory of localization_node.cpp:33-45 behind the yaml `filter:` switch.
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from . import shim
from .params import Params, from_yaml_dict


class FilterChoice(IntEnum):          # filter.h:44-51
    NOT_SET = 0
    EKF_SLAM = 1
    UKF_LOC = 2
    UKF_SLAM = 3
    POSE_GRAPH_SLAM = 4
    NAIVE_COMMAND_PROPAGATION = 5


class Filter:
    """filter.h:54-145.  One filter instance = a batch of size 1 on the GPU."""

    type = FilterChoice.NOT_SET
    _kind = 0

    def __init__(self, max_landmarks: int = 50, max_meas: int = 16, device: int = 0):
        self.isInit = False                    # filter.h:68
        self.map: list[float] = []             # filter.h:69 (localization-only modes; unused by the SLAM filters)
        self.filter_to_compare = FilterChoice.NOT_SET
        self._max_landmarks, self._max_meas, self._device = max_landmarks, max_meas, device
        self._batch: shim.FilterBatch | None = None
        self._params: Params | None = None

    # -- filter.h:59 / ekf.cpp:23-27 / ukf.cpp:25-29
    def readParams(self, config):
        """config: the parsed params.yaml dict (YAML::Node in the reference) or a Params."""
        self._params = config if isinstance(config, Params) else from_yaml_dict(config)
        kind = self._kind
        if self.type == FilterChoice.UKF_LOC:      # localization_node.cpp:36-38 overrides UKF's type before readParams
            kind = shim.UKF_LOC
        self._batch = shim.FilterBatch(kind, self._params.to_c(), 1, self._max_landmarks, self._max_meas,
                                       self._device)

    def setMap(self, landmarks):
        """The /truth/landmarks message (flat float32 [id, x, y]*) the node copies into Filter::map
        (localization_node.cpp: landmark callback; filter.h:68).  Only the localisation-only UKF reads it."""
        self.map = [float(v) for v in np.asarray(landmarks, dtype=np.float32).reshape(-1)]
        if self.type == FilterChoice.UKF_LOC:
            self._need().set_map(np.asarray(self.map, dtype=np.float32))

    def _need(self) -> shim.FilterBatch:
        if self._batch is None:
            raise RuntimeError("readParams must be called before the filter is used.")
        return self._batch

    # -- filter.h:60
    def init(self, x_0: float, y_0: float, yaw_0: float):
        self._need().init(float(np.float32(x_0)), float(np.float32(y_0)), float(np.float32(yaw_0)))
        self.isInit = True

    # -- filter.h:61
    def update(self, cmdMsg, lmMeasMsg):
        """cmdMsg: (fwd, ang) or an object with .fwd/.ang; lmMeasMsg: flat [id,r,b]* or an object with .data."""
        b = self._need()
        fwd, ang = (cmdMsg.fwd, cmdMsg.ang) if hasattr(cmdMsg, "fwd") else cmdMsg
        data = lmMeasMsg.data if hasattr(lmMeasMsg, "data") else lmMeasMsg
        data = np.asarray(data, dtype=np.float32).reshape(-1)
        k = data.size // 3                                    # ekf.cpp:65
        meas, n = b.pack_meas([data[: 3 * k].reshape(k, 3)])
        b.step(fwd, ang, meas, n)
        st = b.status(0)
        if st & shim.STATUS_SAME_STEP_REMATCH:
            # the reference dies here: eigen_assert -> std::runtime_error (filter.h:5, ekf.cpp:115)
            raise RuntimeError("index >= 0 && index < size()")
        if st & shim.STATUS_MEAS_OVERFLOW:
            raise RuntimeError("more detections in one message than max_meas")

    # -- split form (north star): predict from the command, update from the measurements
    def predict(self, fwd: float, ang: float):
        self._need().predict(fwd, ang)

    def correct(self, lmMeasMsg):
        b = self._need()
        data = np.asarray(lmMeasMsg.data if hasattr(lmMeasMsg, "data") else lmMeasMsg, dtype=np.float32).reshape(-1)
        k = data.size // 3
        meas, n = b.pack_meas([data[: 3 * k].reshape(k, 3)])
        b.update(meas, n)

    # -- filter.h:74-76
    def updateNaiveVehPoseEstimate(self, state_vector, landmark_ids):
        raise RuntimeError("updateNaiveVehPoseEstimate is not defined for this filter.")

    def getStateVector(self) -> np.ndarray:
        return self._need().state_vector(0)

    # -- state fields
    @property
    def lm_IDs(self) -> list[int]:             # filter.h:70
        return [int(v) for v in self._need().landmark_ids(0)]

    @property
    def M(self) -> int:
        return self._need().num_landmarks(0)

    @property
    def timestep(self) -> int:
        return self._need().timestep(0)

    @property
    def x_t(self) -> np.ndarray:
        return self._need().state(0)

    @property
    def P_t(self) -> np.ndarray:
        return self._need().cov(0)

    def setupStatePublisher(self, node=None):   # filter.h:65 (ROS plumbing; the topic name is kept for the shim)
        self.state_topic = {FilterChoice.EKF_SLAM: "/state/ekf", FilterChoice.UKF_SLAM: "/state/ukf",
                            FilterChoice.UKF_LOC: "/state/ukf", FilterChoice.NAIVE_COMMAND_PROPAGATION: "/state/naive"}.get(self.type)

    def publishState(self) -> dict:
        raise NotImplementedError


class EKF(Filter):
    """filter.h:148-174, ekf.cpp."""
    type = FilterChoice.EKF_SLAM
    _kind = shim.EKF_SLAM

    def publishState(self) -> dict:
        """EKFState.msg fields exactly as ekf.cpp:192-220 fills them (float32 wire types)."""
        x, P, ids = self.x_t, self.P_t, self.lm_IDs
        lm = np.zeros(3 * len(ids), dtype=np.float32)
        for i, ident in enumerate(ids):
            lm[3 * i] = np.float32(ident)
            lm[3 * i + 1] = x[3 + 2 * i]
            lm[3 * i + 2] = x[4 + 2 * i]
        return dict(timestep=self.timestep, x_v=np.float32(x[0]), y_v=np.float32(x[1]), yaw_v=np.float32(x[2]),
                    M=len(ids), landmarks=lm, P=P.astype(np.float32).reshape(-1))


class UKF(Filter):
    """filter.h:177-223, ukf.cpp (UKF_SLAM; UKF_LOC when `type` is overridden as localization_node.cpp:36-38 does)."""
    type = FilterChoice.UKF_SLAM
    _kind = shim.UKF_SLAM

    def publishState(self) -> dict:
        """UKFState.msg fields as ukf.cpp:60-104 fills them, X included (point-major, ukf.cpp:91-99)."""
        x, P, ids = self.x_t, self.P_t, self.lm_IDs
        lm = np.zeros(3 * len(ids), dtype=np.float32)
        for i, ident in enumerate(ids):
            lm[3 * i] = np.float32(ident)
            lm[3 * i + 1] = x[4 + 2 * i]
            lm[3 * i + 2] = x[5 + 2 * i]
        yaw = np.remainder(np.arctan2(x[3], x[2]) + np.pi, 2 * np.pi) - np.pi
        return dict(timestep=self.timestep, x_v=np.float32(x[0]), y_v=np.float32(x[1]), yaw_v=np.float32(yaw),
                    M=len(ids), landmarks=lm, P=P.astype(np.float32).reshape(-1),
                    X=self._need().sigma_points(0).astype(np.float32).reshape(-1))

    @property
    def X(self) -> np.ndarray:                 # filter.h:199 (n x (2n+1), column j = sigma point j)
        return self._need().sigma_points(0).T.copy()


class NaiveFilter(Filter):
    """filter.h:325-370: ignores the measurements and propagates the pose by the command."""
    type = FilterChoice.NAIVE_COMMAND_PROPAGATION
    _kind = shim.NAIVE

    def publishState(self) -> dict:
        """NaiveState.msg fields as filter.h:358-368 fills them."""
        x = self.x_t
        return dict(timestep=self.timestep, x_v=np.float32(x[0]), y_v=np.float32(x[1]), yaw_v=np.float32(x[2]))


def make_filter(config, **kw) -> Filter:
    """localization_node.cpp:28-47: choose the derived class from `filter:` and read its params."""
    p = config if isinstance(config, Params) else from_yaml_dict(config)
    if p.filter == "ekf_slam":
        f: Filter = EKF(**kw)
    elif p.filter == "ukf_slam":
        f = UKF(**kw)
    elif p.filter == "ukf_loc":
        f = UKF(**kw)
        f.type = FilterChoice.UKF_LOC              # localization_node.cpp:36-38: override the default of UKF_SLAM
    elif p.filter == "pose_graph":
        raise RuntimeError("filter 'pose_graph' is outside the B200 hot path (SURVEY.md section 8f); "
                           "use ekf_slam, ukf_slam or ukf_loc")
    else:
        raise RuntimeError("Invalid filter choice in params.yaml.")   # localization_node.cpp:44
    f.readParams(p)
    return f
