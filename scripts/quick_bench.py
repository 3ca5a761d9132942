"""Scratch timing of the Monte-Carlo sweep (development aid; bench.py is the contract)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from live_ekf_slam_b200 import shim  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    kind = shim.UKF_SLAM if (len(sys.argv) > 3 and sys.argv[3] == "ukf") else shim.EKF_SLAM
    p, lm, fwd, ang = H.config2(seed=0, steps=T)
    fb = shim.FilterBatch(kind, p.to_c(), B, 50, 8)
    sim = shim.Simulator(fb, lm, seed=1)
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    if len(sys.argv) > 5:
        fb.tune(7, int(sys.argv[5]))
    if len(sys.argv) > 6:
        fb.tune(10, int(sys.argv[6]))
    for rep in range(reps):
        fb.reset(0, 0, 0)
        sim.reset()
        fb.set_profiling(rep == reps - 1)
        fb.synchronize()
        t0 = time.time()
        sim.run(fwd, ang)
        fb.synchronize()
        dt = time.time() - t0
        s = fb.stats()
        line = f"rep {rep}: {dt*1e3:.1f} ms  {B*T/dt/1e6:.2f} M updates/s  alg GB/s {s[8]/dt/1e9:.0f}  mean n {s[10]/s[0]:.1f} mean k {s[11]/s[0]:.2f} pos err {s[4]/s[0]:.3f} bad {s[6]}"
        if rep == reps - 1:
            ms, n = fb.profile()
            line += f" | step kernel {ms:.1f} ms over {n} launches -> alg GB/s {s[8]/ms/1e6:.0f}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
