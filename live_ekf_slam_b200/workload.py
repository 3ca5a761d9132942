"""Host-side input preparation for the benchmark configurations: landmark maps and the
precomputed nearest-neighbour TSP command trajectory (sim_node.py:63-152, :155-206).

This is set-up work done once per sweep; the per-step measurement
generator itself runs on the GPU (csrc/sim.cu).  The reference draws from an unseeded
`random.random()`; here a seeded numpy Generator is consumed in the same order.
"""
from __future__ import annotations

import math

import numpy as np

from .params import Params


def grid_map(p: Params) -> np.ndarray:
    """sim_node.py:167-176: square grid, ids row-major over (r, c)."""
    shift = p.map_grid_step / 2
    axis = np.arange(-p.map_bound + shift, p.map_bound, p.map_grid_step)
    return np.array([(r, c) for r in axis for c in axis], dtype=np.float64)


def grid_map_5x10() -> np.ndarray:
    """BASELINE config 2/3 '50-landmark grid' (deviation D-2: the reference generator only makes
    k^2 grids): x in {-8,-4,0,4,8}, y in {-9,-7,...,9}, ids row-major."""
    return np.array([(x, y) for x in (-8.0, -4.0, 0.0, 4.0, 8.0) for y in np.arange(-9.0, 9.5, 2.0)], dtype=np.float64)


def random_map(n: int, bound: float, min_sep: float, rng: np.random.Generator) -> np.ndarray:
    """sim_node.py:177-188 on the blank occupancy map (no obstacle rejection): rejection-sample
    positions at least `min_sep` from every accepted landmark."""
    pts: list[tuple[float, float]] = []
    while len(pts) < n:
        pos = (2 * bound * rng.random() - bound, 2 * bound * rng.random() - bound)
        if any(math.hypot(q[0] - pos[0], q[1] - pos[1]) < min_sep for q in pts):
            continue
        pts.append(pos)
    return np.array(pts, dtype=np.float64)


def random_map_fast(n: int, bound: float, min_sep: float, rng: np.random.Generator) -> np.ndarray:
    """Same acceptance rule as random_map with a cell hash for the separation test (2000-landmark maps)."""
    cell = max(min_sep, 1e-9)
    buckets: dict[tuple[int, int], list[tuple[float, float]]] = {}
    pts = []
    while len(pts) < n:
        pos = (2 * bound * rng.random() - bound, 2 * bound * rng.random() - bound)
        cx, cy = int(math.floor(pos[0] / cell)), int(math.floor(pos[1] / cell))
        ok = True
        for ax in (cx - 1, cx, cx + 1):
            for ay in (cy - 1, cy, cy + 1):
                for q in buckets.get((ax, ay), ()):
                    if math.hypot(q[0] - pos[0], q[1] - pos[1]) < min_sep:
                        ok = False
        if not ok:
            continue
        buckets.setdefault((cx, cy), []).append(pos)
        pts.append(pos)
    return np.array(pts, dtype=np.float64)


def _norm(a, b) -> float:
    return ((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2) ** (1 / 2)


def tsp_trajectory(landmarks: np.ndarray, p: Params, rng: np.random.Generator, num_iterations: int | None = None):
    """sim_node.py:63-152: noisy copy of the map, nearest-neighbour tour, then a command per step
    that drives at most d_max / th_max towards the current goal, rotating the tour when within
    visitation_threshold.  Returns (fwd float32[T], ang float32[T]) -- the float32 wire values of
    Command.msg:3-5."""
    T = p.num_iterations if num_iterations is None else num_iterations
    N = len(landmarks)
    x_t = list(p.init_pose)
    region = (-p.map_bound, p.map_bound)
    noisy = {}
    for i in range(N):
        nx = landmarks[i][0] + 2 * p.landmark_noise * rng.random() - p.landmark_noise
        ny = landmarks[i][1] + 2 * p.landmark_noise * rng.random() - p.landmark_noise
        noisy[i] = (max(region[0] + 1, min(nx, region[1] - 1)), max(region[0] + 1, min(ny, region[1] - 1)))
    cur_goal = 0
    cur_dist = _norm(noisy[0], x_t)
    for i in range(N):
        if _norm(noisy[i], x_t) < cur_dist:
            cur_goal = i
            cur_dist = _norm(noisy[i], x_t)
    cur_node = cur_goal
    path = [cur_node]
    unvisited = [i for i in range(N) if i != cur_node]
    while unvisited:
        cur_dist = -1
        for i in unvisited:
            d = _norm(noisy[i], noisy[cur_node])
            if cur_dist < 0 or d < cur_dist:
                cur_goal, cur_dist = i, d
        path.append(cur_goal)
        cur_node = cur_goal
        unvisited.remove(cur_node)
    fwd, ang = [], []
    for _ in range(T):
        if _norm(x_t, noisy[path[0]]) < p.visitation_threshold:
            path = path[1:] + [path[0]]
        goal = noisy[path[0]]
        d = _norm(goal, x_t)
        gb = math.atan2(goal[1] - x_t[1], goal[0] - x_t[0])
        hdg = math.remainder(gb - x_t[2], math.tau)
        d = min(d, p.d_max)
        if abs(hdg) > p.th_max:
            hdg = p.th_max * float(np.sign(hdg))
        x_t = [x_t[0] + d * math.cos(x_t[2]), x_t[1] + d * math.sin(x_t[2]), x_t[2] + hdg]
        fwd.append(d)
        ang.append(hdg)
    return np.asarray(fwd, dtype=np.float32), np.asarray(ang, dtype=np.float32)
