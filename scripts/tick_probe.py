"""Where does a slam_step_io tick go?  Times T ticks (host sync after each) for a few variants of the path."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from live_ekf_slam_b200 import shim, Params, workload as wl

def run(B, T, variant):
    p = Params(filter="ekf_slam")
    rng = np.random.default_rng(0)
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, rng, T)
    mm = 8
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), B, 50, mm)
    sim = shim.Simulator(fb, lm, seed=1)
    d_fwd, d_ang = torch.from_numpy(fwd).cuda(), torch.from_numpy(ang).cuda()
    h_meas = torch.empty((T, B, mm, 3), dtype=torch.float32).pin_memory()
    h_n = torch.empty((T, B), dtype=torch.int32).pin_memory()
    h_fwd = torch.from_numpy(fwd.copy()).pin_memory(); h_ang = torch.from_numpy(ang.copy()).pin_memory()
    h_pose = torch.empty((T, B, 3), dtype=torch.float64).pin_memory()
    fb.reset(0, 0, 0); sim.reset(0, 0, 0)
    for t in range(T):
        sim.step_device(d_fwd[t:], d_ang[t:], 0, t)
        m, n = sim.meas()
        h_meas[t].copy_(torch.from_numpy(m)); h_n[t].copy_(torch.from_numpy(n))
    fp, ap, mp, npn, pp = h_fwd.data_ptr(), h_ang.data_ptr(), h_meas.data_ptr(), h_n.data_ptr(), h_pose.data_ptr()
    sm_, sn_, sp_ = B * mm * 3 * 4, B * 4, B * 3 * 8
    if "staged" in variant: fb.tune(14, 1)
    if "fullcap" in variant: fb.tune(0, 50)
    poses = "noposes" not in variant
    nosync = "nosync" in variant
    for rep in range(2):
        fb.reset(0, 0, 0); fb.synchronize()
        t0 = time.perf_counter()
        for t in range(T):
            fb.step_io(fp + 4 * t, ap + 4 * t, 0, mp + sm_ * t, npn + sn_ * t, (pp + sp_ * t) if poses else None)
            if not nosync: fb.synchronize()
        fb.synchronize()
        dt = time.perf_counter() - t0
    # device-resident per-step for reference
    fb.reset(0, 0, 0); sim.reset(0, 0, 0); fb.tune(3, 1); fb.synchronize()
    t0 = time.perf_counter(); sim.run_device(d_fwd, d_ang, 0, T, 0); fb.synchronize(); dtd = time.perf_counter() - t0
    print(f"B={B:5d} {variant:28s} tick {1e6*dt/T:7.1f} us  ({B*T/dt/1e6:6.1f} M upd/s)   [device-resident per-step loop incl. sim+err kernels: {1e6*dtd/T:6.1f} us]", flush=True)

for B in (1, 4096):
    for v in ("zero", "zero_noposes", "zero_fullcap", "staged", "staged_fullcap", "zero_nosync", "staged_nosync"):
        run(B, 1000, v)
