"""Multi-GPU plumbing for Monte-Carlo sweeps: one process per GPU, instances sharded contiguously, no data-path
collective; the only exchange is one all-reduce (SUM) of the SLAM_NUM_STATS error/work accumulators
(SURVEY.md section 8e).  torch.distributed is plumbing here (NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import os

import numpy as np


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total_instances: int, rank: int, world: int):
    """Strong-scaling split of a fixed sweep: instance i lives on rank floor(i * world / total).
    Returns (first_global_instance, count)."""
    first = (rank * total_instances + world - 1) // world
    nxt = ((rank + 1) * total_instances + world - 1) // world
    return first, nxt - first


def weak_offset(per_rank_instances: int, rank: int) -> int:
    """Weak scaling: every rank runs `per_rank_instances`; the RNG is keyed by the GLOBAL instance id."""
    return rank * per_rank_instances


def init_distributed(backend: str, device=None):
    import torch.distributed as dist
    rank, local, world = world_info()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, local, world


def allreduce_stats(stats: np.ndarray, device=None) -> np.ndarray:
    """SUM the per-rank statistics vector over all ranks (no-op for a single process)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(stats, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def derive_accuracy(st: np.ndarray, total_instances: int) -> dict:
    """RMSE / mean position error / NEES from the summed accumulators (slam_filter.h SLAM_NUM_STATS layout)."""
    cnt = max(float(st[0]), 1.0)
    return {"rmse_x": float(np.sqrt(st[1] / cnt)), "rmse_y": float(np.sqrt(st[2] / cnt)),
            "rmse_yaw": float(np.sqrt(st[3] / cnt)), "mean_pos_err_m": float(st[4] / cnt),
            "mean_nees3": float(st[5] / cnt), "bad_instances": int(st[6]),
            "mean_final_landmarks": float(st[7] / max(total_instances, 1))}


def allreduce_histogram(counts: np.ndarray, device=None) -> np.ndarray:
    """SUM the per-rank histogram of per-run average errors (exact int64 counts) over all ranks."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int64))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def histogram_summary(counts: np.ndarray, lo: float, hi: float) -> dict:
    """Table-1-style summary of a merged histogram (make_bar_graphs.py:55 prints the mean over runs; at Monte-Carlo scale
    the quantiles are the informative part).  Quantiles are bin upper edges, i.e. exact to one bin width."""
    counts = np.asarray(counts, dtype=np.int64)
    nbins = counts.size - 2
    total = int(counts.sum())
    width = (hi - lo) / nbins
    cum = np.cumsum(counts)
    out = {"runs": total, "below_lo": int(counts[0]), "at_or_above_hi": int(counts[-1]), "bin_width": width}
    for name, q in (("p50", 0.5), ("p90", 0.9), ("p99", 0.99)):
        if total == 0:
            out[name] = None
            continue
        k = int(np.searchsorted(cum, q * total, side="left"))
        out[name] = None if k >= nbins + 1 else float(lo + width * k)     # upper edge of bin k (k = 0: below lo)
    centers = lo + width * (np.arange(nbins) + 0.5)
    inside = counts[1:-1]
    out["mean_binned"] = float((inside * centers).sum() / inside.sum()) if inside.sum() else None
    return out
