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
tot = parallel.allreduce_stats(stats)
print("RESULT " + json.dumps({"rank": rank, "world": world, "stats": tot.tolist(), "finals": finals}), flush=True)
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
