"""Scratch: per-chunk times of the persistent sweep kernel for every CTA width (SLAM_DEBUG_SWEEP=1 prints them to stderr),
and the whole-sweep rate; the final statistics of every run are printed so that bit-identity between variants shows."""
import sys
import time

sys.path.insert(0, ".")
from live_ekf_slam_b200 import shim  # noqa: E402
from tests import helpers as H  # noqa: E402

B, T = 4096, 1000
p, lm, fwd, ang = H.config2(seed=0, steps=T)
fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
sim = shim.Simulator(fb, lm, seed=1)
widths = [int(v) for v in sys.argv[1:]] or [0, 32, 64, 96, 128]
for thr in widths:
    fb.tune(2, thr)
    best = 1e9
    for rep in range(3):
        fb.reset(0, 0, 0); sim.reset(); fb.synchronize()
        t0 = time.time(); sim.run(fwd, ang); fb.synchronize()
        best = min(best, time.time() - t0)
    s = fb.stats()
    print(f"threads {thr:3d}: {best*1e3:7.2f} ms  {B*T/best/1e6:7.2f} M updates/s  stats {s[1]!r} {s[2]!r} {s[4]!r} {s[5]!r}", flush=True)
    print(f"== per-chunk times, threads {thr}", file=sys.stderr, flush=True)
    fb.set_profiling(2)
    fb.reset(0, 0, 0); sim.reset(); sim.run(fwd, ang); fb.synchronize()
    ms, n = fb.profile()
    fb.set_profiling(0)
    print(f"threads {thr:3d}: profiled {ms:.2f} ms over {n} chunk launches", flush=True)
