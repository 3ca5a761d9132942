"""The N > 1 path on CPU: two gloo ranks shard a Monte-Carlo sweep (RNG keyed by the global instance id), run their
instances with the oracle standing in for the device, all-reduce the statistics vector, and must reproduce the
single-process totals bit-for-bit in the instance-local results (shard invariance)."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch.distributed as dist
from live_ekf_slam_b200 import parallel
from oracle import oracle_c as oc
from tests import helpers as H

rank, local, world = parallel.init_distributed("gloo")
TOTAL, T, SEED = 6, 120, 99
p, lm, fwd, ang = H.config2(seed=0, steps=T)
op = H.oracle_params(oc, p)
first, cnt = parallel.shard_range(TOTAL, rank, world)
stats = np.zeros(12)
finals = {}
for g in range(first, first + cnt):
    st, pose, truth, _ = oc.run_instance(oc.EKF_SLAM, op, lm, fwd, ang, SEED, g, 50, oc.STRUCTURED)
    e = pose - truth
    e[:, 2] = np.remainder(e[:, 2] + np.pi, 2 * np.pi) - np.pi
    stats[0] += T; stats[1] += (e[:, 0] ** 2).sum(); stats[2] += (e[:, 1] ** 2).sum(); stats[3] += (e[:, 2] ** 2).sum()
    stats[4] += np.sqrt(e[:, 0] ** 2 + e[:, 1] ** 2).sum()
    finals[g] = pose[-1].tolist()
    avg_errs = globals().setdefault("avg_errs", [])
    avg_errs.append(float(np.sqrt(e[:, 0] ** 2 + e[:, 1] ** 2).sum() / T))       # plotting_node.py:212-214, one number per run
tot = parallel.allreduce_stats(stats)
# per-run average errors -> histogram (the layout of slam_get_error_histogram), merged over the ranks
LO, HI, NB = 0.0, 2.0, 40
counts = np.zeros(NB + 2, dtype=np.int64)
for v in globals().get("avg_errs", []):
    k = 0 if v < LO else (NB + 1 if v >= HI else 1 + min(int((v - LO) / (HI - LO) * NB), NB - 1))
    counts[k] += 1
hist = parallel.allreduce_histogram(counts)
print("RESULT " + json.dumps({"rank": rank, "world": world, "stats": tot.tolist(), "finals": finals,
                              "hist": hist.tolist(), "summary": parallel.histogram_summary(hist, LO, HI)}), flush=True)
if world > 1:
    dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world):
    import json
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), REPO_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for pr in procs:
        out, err = pr.communicate(timeout=300)
        assert pr.returncode == 0, err[-2000:]
        line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
        outs.append(json.loads(line[7:]))
    return outs


def test_two_rank_gloo_sweep_matches_single_process():
    single = _run(1)[0]
    two = _run(2)
    finals = {}
    for o in two:
        finals.update(o["finals"])
        # every rank holds the same all-reduced totals
        np.testing.assert_allclose(o["stats"], two[0]["stats"], rtol=0, atol=0)
    assert sorted(finals) == sorted(single["finals"])
    for k, v in single["finals"].items():
        assert finals[k] == v                      # instance results do not depend on the sharding
    np.testing.assert_allclose(two[0]["stats"], single["stats"], rtol=1e-12)
    assert two[0]["stats"][0] == 6 * 120
    # the merged histogram of per-run average errors is exact and identical on every rank
    assert two[0]["hist"] == two[1]["hist"] == single["hist"] and sum(single["hist"]) == 6
    sm = two[0]["summary"]
    assert sm["runs"] == 6 and sm["p50"] is not None and sm["p50"] <= sm["p90"] <= sm["p99"]
    assert abs(sm["mean_binned"] - single["stats"][4] / single["stats"][0]) <= sm["bin_width"]


def test_histogram_summary_edges():
    import sys as _s
    _s.path.insert(0, ROOT)
    from live_ekf_slam_b200 import parallel
    counts = np.array([1, 0, 5, 3, 1, 2], dtype=np.int64)          # below | 4 bins of width 0.25 | above
    sm = parallel.histogram_summary(counts, 0.0, 1.0)
    assert sm["runs"] == 12 and sm["below_lo"] == 1 and sm["at_or_above_hi"] == 2 and sm["bin_width"] == 0.25
    assert sm["p50"] == 0.5 and sm["p90"] is None                 # the 90th percentile lies in the overflow bin
    assert parallel.histogram_summary(np.zeros(6, dtype=np.int64), 0.0, 1.0)["p50"] is None
