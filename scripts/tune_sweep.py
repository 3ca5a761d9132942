"""Scratch: sweep-kernel tuning grid (chunk length, tile headroom, CTA width) on configs[1]."""
import sys
import time

sys.path.insert(0, ".")
from live_ekf_slam_b200 import shim  # noqa: E402
from tests import helpers as H  # noqa: E402

B, T = 4096, 1000
p, lm, fwd, ang = H.config2(seed=0, steps=T)
fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
sim = shim.Simulator(fb, lm, seed=1)


def run(chunk, head, thr, reps=3):
    fb.tune(5, chunk); fb.tune(6, head); fb.tune(2, thr)
    best = 1e9
    for _ in range(reps):
        fb.reset(0, 0, 0); sim.reset(); fb.synchronize()
        t0 = time.time(); sim.run(fwd, ang); fb.synchronize()
        best = min(best, time.time() - t0)
    print(f"chunk {chunk:4d} headroom {head:2d} threads {thr:3d}: {best*1e3:7.2f} ms  {B*T/best/1e6:7.2f} M updates/s", flush=True)


for chunk in (8, 16, 32, 64, 128, 1000):
    run(chunk, 8, 0)
for head in (2, 4, 6, 12, 50):
    run(32, head, 0)
for thr in (32, 64, 128, 256):
    run(32, 8, thr)
