"""BASELINE config 4: single large-map EKF-SLAM, 2000 landmarks (state up to 4003), unknown association, dense FP64
covariance update on the tensor cores.  Times a steady-state window and reports updates/s plus the achieved FP64
rate of the closing rank-2k contraction (4 k n^2 flops per step).

  python scripts/bench_large.py [n_landmarks=2000] [steps=3000] [window=300]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from live_ekf_slam_b200 import Params, shim, workload as wl  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    Wn = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    p = Params(filter="ekf_slam")
    p.landmark_id_is_known = False
    rng = np.random.default_rng(0)
    t0 = time.time()
    lm = wl.random_map_fast(N, p.map_bound, 0.3, rng)
    fwd, ang = wl.tsp_trajectory(lm, p, rng, T)
    print(f"map+trajectory generated in {time.time()-t0:.1f} s", flush=True)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, N, 128)
    fb.init(0, 0, 0)
    sim = shim.Simulator(fb, lm, seed=1)
    # warm-up / map discovery
    t0 = time.time()
    sim.run(fwd[: T - Wn], ang[: T - Wn], first_step=0)
    fb.synchronize()
    t_disc = time.time() - t0
    s0 = fb.stats()
    M0 = fb.num_landmarks(0)
    fb.set_profiling(True)
    t0 = time.time()
    sim.run(fwd[T - Wn:], ang[T - Wn:], first_step=T - Wn)
    fb.synchronize()
    dt = time.time() - t0
    k_ms, k_n = fb.profile()
    s1 = fb.stats()
    flops = s1[9] - s0[9]
    byts = s1[8] - s0[8]
    out = {"workload": f"single EKF-SLAM, {N} landmarks, unknown IDs, dense map (bound 10, min-sep 0.3)",
           "discovery_steps": T - Wn, "discovery_s": t_disc, "window_steps": Wn, "window_s": dt,
           "updates_per_s": Wn / dt, "landmarks_at_window_start": int(M0), "landmarks_final": int(fb.num_landmarks(0)),
           "mean_n": (s1[10] - s0[10]) / Wn, "mean_k": (s1[11] - s0[11]) / Wn,
           "step_ms_events": k_ms / max(k_n, 1),
           "rank2k_flops_per_step": flops / Wn, "achieved_tflops_whole_step": flops / (k_ms * 1e-3) / 1e12,
           "algorithmic_GBps_whole_step": byts / (k_ms * 1e-3) / 1e9,
           "status": int(fb.status(0)), "pos_err_m": float(np.linalg.norm(fb.poses()[0][:2] - sim.truth()[0][:2]))}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
