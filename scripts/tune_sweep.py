"""Scratch: sweep-kernel tuning grid (chunk length, tile headroom) on configs[1]; SLAM_SWEEP_NO_BALANCE=1 for the unbalanced grid."""
import sys
import time

sys.path.insert(0, ".")
from live_ekf_slam_b200 import shim  # noqa: E402
from tests import helpers as H  # noqa: E402

B, T = 4096, 1000
p, lm, fwd, ang = H.config2(seed=0, steps=T)
fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, 8)
sim = shim.Simulator(fb, lm, seed=1)


def run(chunk, head, thr=0, reps=4):
    fb.tune(5, chunk); fb.tune(6, head); fb.tune(2, thr)
    best = 1e9
    for _ in range(reps):
        fb.reset(0, 0, 0); sim.reset(); fb.synchronize()
        t0 = time.time(); sim.run(fwd, ang); fb.synchronize()
        best = min(best, time.time() - t0)
    print(f"chunk {chunk:4d} headroom {head:2d} threads {thr:3d}: {best*1e3:7.2f} ms  {B*T/best/1e6:7.2f} M updates/s", flush=True)


run(32, 8)
for chunk in (16, 24, 40, 48, 64):
    run(chunk, 8)
for head in (2, 4, 6, 10, 12):
    run(32, head)
for chunk, head in ((16, 4), (24, 4), (24, 6), (48, 10), (16, 6)):
    run(chunk, head)
