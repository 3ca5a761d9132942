"""roofline.traffic: DRAM bytes of ONE launch of a kernel together with the algorithmic and moved-model bytes of that
SAME launch.  Run under ncu with the -k / -s / -c the script prints on stderr when called with --plan:

  python scripts/traffic_probe.py <step|sweep|gemm|ukf> --plan          # prints the ncu selector
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:<kernel> -s <skip> -c 1 --csv --log-file gpurun_out/traffic/<which>.csv python scripts/traffic_probe.py <which>

The workload is deterministic (fixed seeds), so the launch ncu captures is the launch whose statistics the script brackets."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
from live_ekf_slam_b200 import shim, Params, workload as wl

which = sys.argv[1]
plan = "--plan" in sys.argv
os.makedirs("gpurun_out/traffic", exist_ok=True)
B, S = 4096, 900                      # state of the 4096-instance bench after 900 filter steps (n ~ 97)
if which in ("step", "sweep", "ukf"):
    filt = "ukf_slam" if which == "ukf" else "ekf_slam"
    p = Params(filter=filt)
    rng = np.random.default_rng(0)
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, rng, 1000)
    if which == "sweep":
        S = 18 * 48                   # chunk 18 of 48 steps (the library's default chunk length)
        sel = ("ekf_sweep_kernel", None, "chunk 18 (filter steps 864..911) of the 4096-instance sweep (full-capacity tile: one launch)")
    elif which == "step":
        sel = ("ekf_step_kernel", 2 * S, "filter step 900 of the 4096-instance sweep, per-step path (first-pass launch)")
    else:
        S = 300
        sel = ("ukf_(front2|eig3|back3)_kernel", None, "filter step 300 of the 4096-instance UKF sweep")
    if plan:
        print(json.dumps({"kernel_regex": sel[0], "skip": sel[1], "note": sel[2]})); sys.exit(0)
    kind = shim.UKF_SLAM if which == "ukf" else shim.EKF_SLAM
    fb = shim.FilterBatch(kind, p.to_c(), B, 50, 8)
    sim = shim.Simulator(fb, lm, seed=2026, instance_offset=0)
    if which == "step":
        fb.tune(3, 1)
    fb.reset(0, 0, 0); sim.reset(0, 0, 0)
    sim.run(fwd[:S], ang[:S], first_step=0)
    s0 = fb.stats()
    l0 = fb.kernel_launches
    nstep = 48 if which == "sweep" else 1
    sim.run(fwd[S:S + nstep], ang[S:S + nstep], first_step=S)
    fb.synchronize()
    l1 = fb.kernel_launches
    s1 = fb.stats()
else:
    p = Params(filter="ekf_slam"); p.landmark_id_is_known = False
    rng = np.random.default_rng(0)
    lm = wl.random_map_fast(2000, p.map_bound, 0.3, rng)
    S = 2700
    fwd, ang = wl.tsp_trajectory(lm, p, rng, S + 1)
    sel = ("lm_gemm", None, "steady-state step 2700 of the 2000-landmark run (n = 3801)")
    if plan:
        print(json.dumps({"kernel_regex": sel[0], "skip": "count of lm_gemm launches before step 2700 = steps with detections; use --launch-skip from a dry run", "note": sel[2]})); sys.exit(0)
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, 2000, 128)
    fb.init(0, 0, 0)
    sim = shim.Simulator(fb, lm, seed=1)
    sim.run(fwd[:S], ang[:S], first_step=0)
    s0 = fb.stats()
    l0 = fb.kernel_launches
    sim.run(fwd[S:S + 1], ang[S:S + 1], first_step=S)
    fb.synchronize()
    l1 = fb.kernel_launches
    s1 = fb.stats()
out = {"launch": sel[2], "algorithmic_bytes_same_launch": float(s1[8] - s0[8]), "moved_model_bytes_same_launch": float(s1[12] - s0[12]),
       "algorithmic_flops_same_launch": float(s1[9] - s0[9]), "mean_n": float((s1[10] - s0[10]) / max(s1[0] - s0[0], 1)) if which != "gemm" else float(s1[10] - s0[10])}
out["launches_in_bracket"] = int(l1 - l0)
json.dump(out, open(f"gpurun_out/traffic/{which}_probe.json", "w"))
print(json.dumps(out))
