#!/usr/bin/env python
"""bench.py -- benchmark of the filter hot path (contract: the task statement / DESIGN.md section 7).

Metric (BASELINE.json): EKF/UKF-SLAM filter updates/sec (batched instances x steps).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle, dense-faithful) on host cores

ONE JSON line.  Its top level is BASELINE configs[1] (4096 Monte-Carlo EKF-SLAM instances per GPU, 50-landmark 5x10 grid,
1000 filter steps, known IDs); a bench "step" is one whole sweep = instances x filter_steps Filter::update() calls:
  value     whole-job updates/s with commands, map and filter state resident in HBM (on-GPU simulator -> filter);
  e2e       the same sweep through the C-ABI with HOST buffers (pinned), copies inside the timed region;
  roofline  the dominant kernel against the measured peak, with the streaming model, the bytes really moved and the real
            limiter spelled out;
  cpu_baseline  the dense-faithful oracle on the host cores (bounded sample).
`configs` carries the other BASELINE configurations as bounded sub-records, each with its own value / ms_per_step / roofline /
cpu_baseline / e2e:  configs.ukf (configs[2]: 4096 UKF-SLAM instances), configs.large (configs[3]: one 2000-landmark EKF,
unknown IDs, DMMA contraction), configs.mixed (configs[4]: 65 536 mixed EKF + UKF instances, STRONG split over the ranks with
the statistics / histogram all-reduce inside the timed region).  `--filter ekf|ukf|large|mixed` runs one of them alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EKF/UKF-SLAM filter updates/sec (batched instances x steps)"
UNIT = "updates/s"


def build_workload(filt: str, T: int):
    from live_ekf_slam_b200 import Params, workload as wl
    p = Params(filter=filt)
    rng = np.random.default_rng(0)
    lm = wl.grid_map_5x10()
    fwd, ang = wl.tsp_trajectory(lm, p, rng, T)
    return p, lm, fwd, ang


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """dram bytes of ONE launch from a committed `ncu --set full` capture together with the algorithmic and moved-model bytes
    of that SAME launch (profiles/ncu_traffic.json, written by scripts/traffic_probe.py); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(kernel)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, enabled: bool = True):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t = None
        self.enabled = enabled

    def start(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return self

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.t = threading.Thread(target=pump, daemon=True)
        self.t.start()
        return self

    def stop(self):
        if not self.enabled:
            return None
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: the oracle in dense-faithful mode
    (the same O(n^3) products ekf.cpp:61,140,172 execute; Eigen/ROS cannot be built here), one filter
    instance per core, all cores busy."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle_c as oc
    from tests import helpers as H
    filt = "ukf" if args.filter == "ukf" else "ekf"
    T = args.filter_steps
    p, lm, fwd, ang = build_workload("ekf_slam" if filt == "ekf" else "ukf_slam", T)
    op = H.oracle_params(oc, p)
    kind = oc.EKF_SLAM if filt == "ekf" else oc.UKF_SLAM
    cores = os.cpu_count() or 1
    per = args.ref_instances_per_core or (16 if filt == "ekf" else 8)
    for _ in range(args.warmup):
        oc.bench(kind, op, lm, fwd[: max(10, T // 20)], ang[: max(10, T // 20)], 0, cores, 1, 50, oc.DENSE)
    tot_s, tot_u = 0.0, 0
    for _ in range(args.steps):
        s, u = oc.bench(kind, op, lm, fwd, ang, 0, cores, per, 50, oc.DENSE)
        tot_s += s; tot_u += u
    val = tot_u / tot_s
    sample = f"{cores * per} instances x {T} steps per bench step ({cores} threads x {per}), dense-faithful oracle (gcc -O2)"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": batch_config(filt, cores * per, T, args.gpus),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0


def batch_config(filt, instances, T, gpus):
    if filt == "ekf":
        wlk = (f"{instances} Monte-Carlo EKF-SLAM instances per GPU, 50-landmark 5x10 grid map, {T} filter steps per sweep, "
               "shared TSP command trajectory, known IDs (BASELINE configs[1])")
    else:
        wlk = (f"{instances} UKF-SLAM instances per GPU, 50 landmarks (state <= 104, 209 sigma points), {T} steps "
               "(BASELINE configs[2])")
    return {"workload": wlk, "instances_per_gpu": instances, "filter_steps": T, "landmarks": 50,
            "l2": "no explicit flush: the covariance working set grows to instances x 16 n^2 B = 350 MB (> 126 MB L2) "
                  "and every step rewrites all of it",
            "parallelism": f"instances sharded over {gpus} GPU(s), no data-path collective; one all-reduce of error stats"}


_FP64 = {}


def fp64_ceiling(torch, n: int = 4096, reps: int = 5):
    """Measured FP64 ceiling of this GPU: cuBLAS DGEMM (n^3) through torch.matmul, CUDA events (FP64 peak is not
    in MEASURED_PEAKS.json, so the bench measures one and prints it next to the nominal figure).  Measured once per process."""
    if "v" in _FP64:
        return _FP64["v"]
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    _FP64["v"] = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b
    torch.cuda.empty_cache()
    return _FP64["v"]


FP64_SRC = ("measured here: cuBLAS DGEMM 4096^3 through torch.matmul (FP64 pipe / DMMA ceiling; nominal B200 FP64 "
            "~37-40 TFLOP/s)")


class Ctx:
    """process-wide plumbing shared by the records"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from live_ekf_slam_b200 import shim, parallel
        self.torch, self.dist, self.shim, self.parallel, self.args = torch, dist, shim, parallel, args
        self.rank, self.local, self.world = parallel.world_info()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        parallel.init_distributed("nccl", self.dev)
        shim.load()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


# ------------------------------------------------------------------------------------------------ batched EKF / UKF record
def measure_batch(cx: Ctx, filt: str, B: int, T: int, K: int, W: int, warm_T: int, e2e_sweeps: int, per_tick_T: int,
                  cpu_per_core: int, with_cpu: bool, with_hist: bool, knobs=()):
    """One Monte-Carlo configuration (BASELINE configs[1] or configs[2]) on this rank's GPU, weak scaling over ranks.
    W warm-up sweeps of warm_T filter steps, then exactly K timed sweeps of T filter steps."""
    torch, shim, parallel, args = cx.torch, cx.shim, cx.parallel, cx.args
    kind = shim.EKF_SLAM if filt == "ekf" else shim.UKF_SLAM
    p, lm, fwd, ang = build_workload("ekf_slam" if filt == "ekf" else "ukf_slam", T)
    fb = shim.FilterBatch(kind, p.to_c(), B, 50, args.max_meas, device=cx.local)
    sim = shim.Simulator(fb, lm, seed=args.seed, instance_offset=parallel.weak_offset(B, cx.rank))   # RNG keyed by the GLOBAL instance id
    for k, v in knobs:
        fb.tune(k, v)
    stream = torch.cuda.ExternalStream(fb.stream, device=cx.dev)
    d_fwd = torch.from_numpy(fwd).cuda()
    d_ang = torch.from_numpy(ang).cuda()
    torch.cuda.synchronize()

    def sweep(steps=T):
        fb.reset(*p.init_pose)
        sim.reset(*p.init_pose)
        sim.run_device(d_fwd, d_ang, 0, steps, 0)

    for _ in range(W):
        sweep(warm_T)
    cx.barrier()
    launches0 = fb.kernel_launches
    clocks = ClockSampler(cx.local, cx.rank == 0).start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_in_timed = filt == "ukf"        # the UKF step is three launches per update: per-launch events cost nothing beside them
    if prof_in_timed:
        fb.set_profiling(1)
    cx.barrier()
    ev0.record(stream)
    for _ in range(K):
        sweep()
    ev1.record(stream)
    fb.synchronize()
    cx.barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = fb.kernel_launches - launches0
    (ms,) = cx.max_over_ranks(ms)
    value = cx.world * B * T * K / (ms * 1e-3)

    # ---- accuracy statistics of the last sweep, summed over ranks (the only collective on this path)
    loc_value = fb.stats()
    st = parallel.allreduce_stats(loc_value, cx.dev)                        # NCCL over NVLink when world > 1
    acc = parallel.derive_accuracy(st, cx.world * B)
    if filt == "ukf":
        r = fb.ukf_routes()
        acc["ukf_route_instance_steps"] = {"dense_gen3": int(r[0]), "ql_gen2": int(r[1]), "explicit_gen1_or_rescue": int(r[2])}
    if with_hist:
        # per-run average position error (the reference's one number per run, plotting_node.py:195-218): on-device histogram,
        # exact counts merged over the ranks, quantiles read off the merged histogram (1 cm bins)
        run_hist = parallel.allreduce_histogram(fb.error_histogram(0.0, 10.0, 1000), cx.dev)
        acc["per_run_avg_pos_err_m"] = parallel.histogram_summary(run_hist, 0.0, 10.0)

    # ---- roofline: CUDA events on the launching stream around the kernel launches
    peak, peak_src = measured_peaks()
    if prof_in_timed:
        k_ms, k_n = fb.profile()
        fb.set_profiling(0)
        loc = loc_value
        k_ms, k_n = k_ms / K, k_n / K
    else:
        fb.tune(3, 1)                      # per-step launches (the per-call C-ABI path's kernel)
        fb.set_profiling(1)
        sweep()
        k_ms, k_n = fb.profile()
        fb.set_profiling(0)
        fb.tune(3, 0)
        loc = fb.stats()
    n_upd = max(loc[0], 1)
    step_roof = {"kernel": "ekf_step_kernel" if filt == "ekf" else "ukf step (ukf_front2_kernel + ukf_eig3_kernel + ukf_back3_kernel; the size-class and hand-over launches of each included)",
                 "kernel_ms_per_launch": k_ms / max(k_n, 1), "launches_timed": int(k_n),
                 "mean_n": loc[10] / n_upd, "mean_k": loc[11] / n_upd,
                 "algorithmic_bytes_per_launch": loc[8] / max(k_n, 1), "moved_bytes_per_launch_model": loc[12] / max(k_n, 1),
                 "streaming_equivalent_gbs": loc[8] / (k_ms * 1e-3) / 1e9, "moved_model_gbs": loc[12] / (k_ms * 1e-3) / 1e9,
                 "algorithmic_tflops": loc[9] / (k_ms * 1e-3) / 1e12, "executed_model_tflops": loc[13] / (k_ms * 1e-3) / 1e12}
    ceil_tf = fp64_ceiling(torch)
    if filt == "ekf":
        # (a) the per-step kernel: P crosses HBM once each way per step -- PACKED (lower triangle), so it really moves about
        #     half of the streaming model's 16 n^2 bytes
        tr = ncu_traffic("ekf_step_kernel")
        step_roof.update({"bound": "hbm", "achieved": step_roof["streaming_equivalent_gbs"], "peak": peak, "unit": "GB/s",
                          "frac": step_roof["streaming_equivalent_gbs"] / peak, "peak_source": peak_src,
                          "frac_meaning": "contract figure: ALGORITHMIC bytes (SURVEY 8d streaming model, 16 n^2 + ...) / kernel time / "
                                          "measured copy peak; the kernel stores P packed-symmetric and really moves "
                                          "`moved_model_gbs` (hbm_utilisation_model = that / peak)",
                          "hbm_utilisation_model": step_roof["moved_model_gbs"] / peak,
                          "traffic": tr["dram_bytes"] if tr else None, "traffic_same_launch": tr})
        # (b) the persistent sweep kernel on the `value` path: P stays in shared memory for a chunk of steps
        fb.set_profiling(2)
        sweep()
        s_ms, s_n = fb.profile()
        fb.set_profiling(0)
        loc2 = fb.stats()
        ach = loc2[8] / (s_ms * 1e-3) / 1e9 if s_ms > 0 else 0.0
        tr2 = ncu_traffic("ekf_sweep_kernel")
        roofline = {"bound": "hbm", "kernel": "ekf_sweep_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "peak_source": peak_src,
                    "real_limiter": "NOT HBM: P is resident in shared memory for the steps of a launch (DRAM busy ~2 % in the ncu "
                                    "capture). One instance-step is a ~6.5 us chain of dependent scalar FP64 work (sincos, sqrt, "
                                    "divisions, atan2), so throughput = resident instances per SM / that latency (10 one-warp CTAs per "
                                    "SM on small tiles .. 4 four-warp CTAs at the full map) until the shared-memory pipe saturates "
                                    "(ncu at the 50-landmark tile: l1tex shared-memory wavefronts 74 % of peak, issue slots 46 %, FP64 "
                                    "pipe 18 %). `frac` is the contract's streaming-equivalent figure (what a per-step streaming design "
                                    "would have had to move), it can exceed 1 and is not an HBM utilisation",
                    "moved_model_gbs": loc2[12] / (s_ms * 1e-3) / 1e9 if s_ms > 0 else 0.0,
                    "hbm_utilisation_model": loc2[12] / (s_ms * 1e-3) / 1e9 / peak if s_ms > 0 else 0.0,
                    "fp64_frac": {"executed_model_tflops": loc2[13] / (s_ms * 1e-3) / 1e12, "algorithmic_tflops": loc2[9] / (s_ms * 1e-3) / 1e12,
                                  "ceiling_tflops": ceil_tf, "frac_executed": loc2[13] / (s_ms * 1e-3) / 1e12 / ceil_tf,
                                  "ceiling_source": FP64_SRC},
                    "traffic": tr2["dram_bytes"] if tr2 else None, "traffic_same_launch": tr2,
                    "algorithmic_bytes_per_launch": loc2[8] / max(s_n, 1), "moved_bytes_per_launch_model": loc2[12] / max(s_n, 1),
                    "kernel_ms_per_launch": s_ms / max(s_n, 1), "launches_timed": int(s_n),
                    "kernel_share_of_sweep": s_ms / (ms / K) if ms > 0 else None,
                    "mean_n": loc2[10] / max(loc2[0], 1), "mean_k": loc2[11] / max(loc2[0], 1),
                    "step_kernel": step_roof}
    else:
        # the UKF step is FP64-compute bound (AI ~ n flop/B): nominal flops against the measured FP64 ceiling
        ach_tf = step_roof["algorithmic_tflops"]
        roofline = {"bound": "tensor", "kernel": step_roof["kernel"], "achieved": ach_tf, "peak": ceil_tf, "unit": "TFLOP/s",
                    "frac": ach_tf / ceil_tf if ceil_tf > 0 else None, "peak_source": FP64_SRC,
                    "frac_meaning": "nominal flops of SURVEY 8d per update (9 n^3 eigh + 2 n^3 sqrt + 2 n^2 (2n+1) contraction + "
                                    "12 k n^2) / kernel time / measured DGEMM ceiling: a throughput-equivalent rate; the kernels "
                                    "execute `executed_model_tflops` (no explicit eigenvectors, no dense sigma matrices)",
                    "executed_model_tflops": step_roof["executed_model_tflops"],
                    "frac_executed": step_roof["executed_model_tflops"] / ceil_tf if ceil_tf > 0 else None,
                    "traffic": None, "kernel_ms_per_launch": step_roof["kernel_ms_per_launch"],
                    "launches_timed": step_roof["launches_timed"], "mean_n": step_roof["mean_n"], "mean_k": step_roof["mean_k"],
                    "kernel_share_of_sweep": k_ms / (ms / K) if ms > 0 else None,
                    "hbm_view": {"streaming_equivalent_gbs": step_roof["streaming_equivalent_gbs"],
                                 "moved_model_gbs": step_roof["moved_model_gbs"], "peak": peak}}

    # ---- e2e: the C-ABI with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if e2e_sweeps > 0:
        mm = args.max_meas
        h_meas = torch.empty((T, B, mm, 3), dtype=torch.float32).pin_memory()
        h_n = torch.empty((T, B), dtype=torch.int32).pin_memory()
        h_fwd = torch.from_numpy(fwd.copy()).pin_memory()
        h_ang = torch.from_numpy(ang.copy()).pin_memory()
        h_pose = torch.empty((T, B, 3), dtype=torch.float64).pin_memory()
        fb.reset(*p.init_pose); sim.reset(*p.init_pose)
        for t in range(T):   # record the message stream once (untimed): this is what a host-side caller would hold
            sim.step_device(d_fwd[t:], d_ang[t:], 0, t)
            m, n = sim.meas()
            h_meas[t].copy_(torch.from_numpy(m)); h_n[t].copy_(torch.from_numpy(n))
        fp, ap, mp, npn, pp = h_fwd.data_ptr(), h_ang.data_ptr(), h_meas.data_ptr(), h_n.data_ptr(), h_pose.data_ptr()
        sm_, sn_, sp_ = B * mm * 3 * 4, B * 4, B * 3 * 8

        def e2e_replay():
            # the public call for a recorded run: slam_run_io (HOST buffers in, HOST poses out; chunks are uploaded,
            # filtered and downloaded in a pipeline inside the library)
            fb.reset(*p.init_pose)
            fb.run_io(fp, ap, 0, mp, npn, pp, T)
            fb.synchronize()

        def e2e_ticks(n_ticks):
            # one slam_step_io per reference timer tick, host sync after every tick (poses readable each tick):
            # the Filter::update contract of localization_node.cpp:108-131
            fb.reset(*p.init_pose)
            for t in range(n_ticks):
                fb.step_io(fp + 4 * t, ap + 4 * t, 0, mp + sm_ * t, npn + sn_ * t, pp + sp_ * t)
                fb.synchronize()

        if filt == "ekf":
            e2e_replay()   # warm-up (the UKF sub-record is bounded to one replay: its kernels are warm from the value path)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_sweeps):
            e2e_replay()
        cx.barrier()
        dt_replay = time.perf_counter() - t0
        h_pose_replay = h_pose.clone()
        nt = min(per_tick_T, T)
        wn = nt if filt == "ekf" else min(nt, 50)
        e2e_ticks(wn)      # warm-up of the per-tick path; also cross-checks the two paths
        tick_vs_replay = float((h_pose[:wn] - h_pose_replay[:wn]).abs().max())
        cx.barrier()
        t0 = time.perf_counter()
        e2e_ticks(nt)
        cx.barrier()
        dt_tick = time.perf_counter() - t0
        # the same ticks without a host sync in between (stream-ordered; the poses of every tick still land in the host buffer):
        # what a caller gets who reads the estimates a few ticks late
        cx.barrier()
        t0 = time.perf_counter()
        fb.reset(*p.init_pose)
        for t in range(nt):
            fb.step_io(fp + 4 * t, ap + 4 * t, 0, mp + sm_ * t, npn + sn_ * t, pp + sp_ * t)
        fb.synchronize()
        cx.barrier()
        dt_async = time.perf_counter() - t0
        async_vs_tick = float((h_pose[:wn] - h_pose_replay[:wn]).abs().max())
        dt_replay, dt_tick, dt_async = cx.max_over_ranks(dt_replay, dt_tick, dt_async)
        e2e = {"value": cx.world * B * T * e2e_sweeps / dt_replay, "unit": UNIT,
               "h2d_bytes_per_step": T * (8 + sm_ + sn_), "d2h_bytes_per_step": T * sp_,
               "mode": "slam_run_io: the recorded run (commands + [id,r,b] messages of all T filter steps) in pinned "
                       "HOST memory -> pose estimates of every filter step in pinned HOST memory; host wall clock, "
                       "copies inside the timed region",
               "sweeps": e2e_sweeps,
               "per_tick_value": cx.world * B * nt / dt_tick,
               "per_tick_mode": f"slam_step_io per filter step with a host sync after every step ({nt} ticks timed): the "
                                "Filter::update call a ROS node makes",
               "per_tick_us": 1e6 * dt_tick / nt,
               "per_tick_path": "pinned host buffers are read / written in place by the kernels (zero copy); batched EKF: one launch per tick",
               "per_tick_async_value": cx.world * B * nt / dt_async,
               "per_tick_async_mode": "the same slam_step_io ticks issued back to back, one host sync at the end",
               "per_tick_async_max_pose_diff": async_vs_tick,
               "per_tick_vs_replay_max_pose_diff": tick_vs_replay}
        del h_meas, h_n, h_pose

    # ---- CPU baseline (rank 0, N=1 only): dense-faithful oracle, one instance per core, bounded sample
    cpu = None
    if with_cpu and cx.rank == 0 and cx.world == 1:
        from oracle import oracle_c as oc
        from tests import helpers as H
        op = H.oracle_params(oc, p)
        okind = oc.EKF_SLAM if filt == "ekf" else oc.UKF_SLAM
        cores = os.cpu_count() or 1
        s, u = oc.bench(okind, op, lm, fwd, ang, 0, cores, cpu_per_core, 50, oc.DENSE)
        cpu = {"value": u / s, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores * cpu_per_core} instances x {T} steps ({cores} threads x {cpu_per_core}), dense-faithful oracle, {s:.1f} s"}

    rec = {"value": value, "unit": UNIT, "steps": K, "warmup": W, "ms_per_step": ms / K, "scaling": "weak", "dtype": "f64",
           "config": batch_config(filt, B, T, cx.world), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
           "roofline": roofline, "cpu_baseline": cpu, "accuracy": acc}
    if warm_T != T:
        rec["warmup_note"] = f"{W} warm-up sweeps of {warm_T} filter steps each, then {K} timed sweep(s) of {T} (bounded sub-record)"
    sim.close(); fb.close()
    del d_fwd, d_ang
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------ large map (configs[3])
def measure_large(cx: Ctx, N: int, T_disc: int, Wn: int, T_e2e: int, cpu_steps: int, with_cpu: bool):
    """BASELINE configs[3]: ONE EKF-SLAM instance, 2000 landmarks on the dense map (bound 10, generation min-sep 0.3), unknown-ID
    box-gate association.  The map is discovered first (untimed: T_disc steps, also the warm-up), then a steady-state window of Wn
    filter steps is timed.  Does not shard: every rank runs an independent replica."""
    torch, shim, args = cx.torch, cx.shim, cx.args
    from live_ekf_slam_b200 import Params, workload as wl
    p = Params(filter="ekf_slam")
    p.landmark_id_is_known = False
    rng = np.random.default_rng(0)
    lm = wl.random_map_fast(N, p.map_bound, 0.3, rng)
    T = T_disc + 3 * Wn + T_e2e
    fwd, ang = wl.tsp_trajectory(lm, p, rng, T)
    mm = 128
    fb = shim.FilterBatch(shim.EKF_SLAM, p.to_c(), 1, N, mm, device=cx.local)
    fb.init(0, 0, 0)
    sim = shim.Simulator(fb, lm, seed=1 + cx.rank)
    stream = torch.cuda.ExternalStream(fb.stream, device=cx.dev)
    sim.run(fwd[:T_disc], ang[:T_disc], first_step=0)                       # discovery = warm-up (thousands of steps)
    fb.synchronize()
    M0 = fb.num_landmarks(0)
    s0 = fb.stats()
    cx.barrier()
    clocks = ClockSampler(cx.local, cx.rank == 0).start()
    l0 = fb.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = T_disc
    ev0.record(stream)
    sim.run(fwd[t:t + Wn], ang[t:t + Wn], first_step=t)
    ev1.record(stream)
    fb.synchronize()
    cx.barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = fb.kernel_launches - l0
    s1 = fb.stats()
    (ms,) = cx.max_over_ranks(ms)
    value = cx.world * Wn / (ms * 1e-3)
    # per-kernel timing, two more windows: whole filter step (predict + front + gemm + commit), then lm_gemm alone
    t += Wn
    fb.set_profiling(1)
    sim.run(fwd[t:t + Wn], ang[t:t + Wn], first_step=t)
    step_ms, step_n = fb.profile()
    sa = fb.stats()
    t += Wn
    fb.set_profiling(3)
    sim.run(fwd[t:t + Wn], ang[t:t + Wn], first_step=t)
    gemm_ms, gemm_n = fb.profile()
    fb.set_profiling(0)
    sb = fb.stats()
    t += Wn
    ceil_tf = fp64_ceiling(torch)
    fl_step, fl_gemm = sa[9] - s1[9], sb[9] - sa[9]                        # 4 k n^2 per step
    gemm_tf = fl_gemm / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_tf = fl_step / (step_ms * 1e-3) / 1e12 if step_ms > 0 else 0.0
    tr = ncu_traffic("lm_gemm")
    roofline = {"bound": "tensor", "kernel": "lm_gemm (rank-2k FP64 DMMA contraction P += U G)", "achieved": gemm_tf, "peak": ceil_tf,
                "unit": "TFLOP/s", "frac": gemm_tf / ceil_tf if ceil_tf > 0 else None, "peak_source": FP64_SRC,
                "frac_meaning": "algorithmic flops 4 k n^2 per step (SURVEY 8d: the k rank-2 updates as one rank-2k GEMM over the FULL "
                                "square) / lm_gemm time / measured DGEMM ceiling",
                "executed_tflops": (sb[13] - sa[13]) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0,
                "kernel_ms_per_launch": gemm_ms / max(gemm_n, 1), "launches_timed": int(gemm_n),
                "whole_step": {"achieved": step_tf, "frac": step_tf / ceil_tf if ceil_tf > 0 else None,
                               "ms_per_step": step_ms / max(step_n, 1), "launches_timed": int(step_n),
                               "kernels": "lm_predict_rows + lm_predict_cols + lm_front + lm_gemm + lm_commit"},
                "gemm_share_of_step": (gemm_ms / max(gemm_n, 1)) / (step_ms / max(step_n, 1)) if step_ms > 0 and step_n else None,
                "traffic": tr["dram_bytes"] if tr else None, "traffic_same_launch": tr,
                "mean_n": (s1[10] - s0[10]) / Wn, "mean_k": (s1[11] - s0[11]) / Wn,
                "hbm_view": {"streaming_equivalent_gbs": (s1[8] - s0[8]) / (ms * 1e-3) / 1e9}}
    # ---- e2e: Filter::update per tick through HOST buffers (pinned), pose read back every tick
    e2e, cpu = None, None
    x_snap = P_snap = ids_snap = None
    msgs = []
    if T_e2e > 0:
        h_meas = torch.zeros((T_e2e, mm, 3), dtype=torch.float32).pin_memory()
        h_n = torch.zeros((T_e2e,), dtype=torch.int32).pin_memory()
        h_fwd = torch.from_numpy(fwd[t:t + T_e2e].copy()).pin_memory()
        h_ang = torch.from_numpy(ang[t:t + T_e2e].copy()).pin_memory()
        h_pose = torch.zeros((T_e2e, 3), dtype=torch.float64).pin_memory()
        for q in range(T_e2e):          # the simulator alone runs ahead (it does not depend on the filter): record the messages
            sim.step(fwd[t + q], ang[t + q], t + q)
            m, n = sim.meas()
            h_meas[q].copy_(torch.from_numpy(m[0])); h_n[q] = int(n[0])
            if q < cpu_steps:
                msgs.append(m[0, : n[0]].copy())
        if with_cpu and cx.rank == 0 and cx.world == 1 and cpu_steps > 0:
            x_snap, P_snap, ids_snap, ts_snap = fb.state(0), fb.cov(0), fb.landmark_ids(0), fb.timestep(0)
        fp, ap, mp, npn, pp = h_fwd.data_ptr(), h_ang.data_ptr(), h_meas.data_ptr(), h_n.data_ptr(), h_pose.data_ptr()
        cx.barrier()
        t0 = time.perf_counter()
        for q in range(T_e2e):
            fb.step_io(fp + 4 * q, ap + 4 * q, 0, mp + mm * 12 * q, npn + 4 * q, pp + 24 * q)
            fb.synchronize()
        cx.barrier()
        dt = time.perf_counter() - t0
        (dt,) = cx.max_over_ranks(dt)
        e2e = {"value": cx.world * T_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": 8 + mm * 12 + 4, "d2h_bytes_per_step": 24,
               "mode": f"slam_step_io per filter step ({T_e2e} steady-state ticks): command + [id,r,b] message from pinned HOST memory, "
                       "pose estimate back to pinned HOST memory, host sync every tick; host wall clock",
               "final_pose": [float(v) for v in h_pose[-1]]}
    if x_snap is not None:
        # CPU baseline: the oracle from the same committed state on the same messages.  Dense-faithful mode is out of reach of a
        # bounded sample at this size (ekf.cpp:140 is a dense n^3 product per landmark update: 2 n^3 = 1.1e11 flop x ~70 updates per
        # step), so the STRUCTURED mode is timed (same values, the exactly-zero terms skipped: O(k n^2) per step) -- this flatters
        # the reference by orders of magnitude and is stated as such.
        from oracle import oracle_c as oc
        from tests import helpers as H
        of = oc.OracleFilter(oc.EKF_SLAM, H.oracle_params(oc, p), N)
        of.set_state(x_snap, P_snap, ids_snap, ts_snap)
        t0 = time.perf_counter()
        for q in range(cpu_steps):
            of.update(fwd[t + q], ang[t + q], msgs[q], oc.STRUCTURED)
        dtc = time.perf_counter() - t0
        nd = float(len(x_snap))
        kk = float(np.mean([len(m) for m in msgs]))
        cpu = {"value": cpu_steps / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{cpu_steps} steady-state steps (n = {int(nd)}, k ~ {kk:.0f}) from the GPU's committed state, oracle in STRUCTURED "
                         f"mode (skips the exactly-zero terms of the reference's dense products), {dtc:.1f} s",
               "dense_faithful_note": f"the reference's literal products cost ~{2 * nd ** 3 * kk + 4 * nd ** 3:.2e} flop per step at this size "
                                      "(ekf.cpp:61,140): hours per step on one core, not measurable inside a bounded sample"}
        del P_snap
    rec = {"value": value, "unit": UNIT, "steps": 1, "warmup": 1, "ms_per_step": ms, "filter_steps_timed": Wn,
           "ms_per_filter_step": ms / Wn, "scaling": "replicas only (one independent replica per GPU)", "dtype": "f64",
           "config": {"workload": f"single large-map EKF-SLAM, {N} landmarks on the dense map (bound 10, min-sep 0.3), unknown-ID box-gate "
                                  f"association, on-GPU simulator (BASELINE configs[3]); {T_disc} discovery steps (untimed warm-up), then a "
                                  f"steady-state window of {Wn} filter steps",
                      "landmarks_at_window_start": int(M0), "state_dim": 3 + 2 * int(M0), "max_meas": mm,
                      "l2": "P is 3 + 2M squared doubles = 116 MB at M = 1900: about the size of the 126 MB L2; every step rewrites all of it"},
           "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "accuracy": {"status": int(fb.status(0)), "landmarks_final": int(fb.num_landmarks(0)),
                        "pos_err_m": float(np.linalg.norm(fb.poses()[0][:2] - sim.truth()[0][:2]))}}
    sim.close(); fb.close()
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------ mixed sweep (configs[4])
def measure_mixed(cx: Ctx, total: int, T: int, W: int, warm_T: int):
    """BASELINE configs[4]: `total` instances, half EKF-SLAM and half UKF-SLAM, STRONG split over the ranks: each kind's global
    instance range is cut with parallel.shard_range (so every rank holds the same share of cheap EKF and expensive UKF instances --
    a split by global index alone would put all UKF instances, ~200x the cost, on half of the GPUs).  The two kinds run
    concurrently on their handles' streams.  The timed region ends with the all-reduce of the error statistics and of the per-run
    error histograms of both kinds (NCCL over NVLink when world > 1)."""
    torch, shim, parallel, args = cx.torch, cx.shim, cx.parallel, cx.args
    half = total // 2
    arms = []
    for name, kind, filt, goff in (("ekf", shim.EKF_SLAM, "ekf_slam", 0), ("ukf", shim.UKF_SLAM, "ukf_slam", half)):
        p, lm, fwd, ang = build_workload(filt, T)
        first, cnt = parallel.shard_range(half, cx.rank, cx.world)
        fb = shim.FilterBatch(kind, p.to_c(), cnt, 50, args.max_meas, device=cx.local)
        sim = shim.Simulator(fb, lm, seed=args.seed, instance_offset=goff + first)     # global instance id: EKF first, then UKF
        st = torch.cuda.ExternalStream(fb.stream, device=cx.dev)
        arms.append({"name": name, "fb": fb, "sim": sim, "p": p, "stream": st, "count": cnt,
                     "fwd": torch.from_numpy(fwd).cuda(), "ang": torch.from_numpy(ang).cuda()})
    torch.cuda.synchronize()

    def sweep(steps):
        for a in arms:
            a["fb"].reset(*a["p"].init_pose)
            a["sim"].reset(*a["p"].init_pose)
            a["sim"].run_device(a["fwd"], a["ang"], 0, steps, 0)

    def reduce_all():
        out = {}
        for a in arms:
            stt = parallel.allreduce_stats(a["fb"].stats(), cx.dev)
            hist = parallel.allreduce_histogram(a["fb"].error_histogram(0.0, 10.0, 1000), cx.dev)
            out[a["name"]] = (stt, hist)
        return out

    for _ in range(W):
        sweep(warm_T)
    reduce_all()                                         # warm the collective too
    cx.barrier()
    l0 = sum(a["fb"].kernel_launches for a in arms)
    clocks = ClockSampler(cx.local, cx.rank == 0).start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in arms]
    cx.barrier()
    t0 = time.perf_counter()
    for a, (e0, _) in zip(arms, evs):
        e0.record(a["stream"])
    sweep(T)
    for a, (_, e1) in zip(arms, evs):
        e1.record(a["stream"])
    red = reduce_all()                                   # synchronises both arms, then the collectives: inside the timed region
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cx.barrier()
    clk = clocks.stop()
    per_arm = [e0.elapsed_time(e1) for (e0, e1) in evs]
    launches = sum(a["fb"].kernel_launches for a in arms) - l0
    dt, arm_ekf, arm_ukf = cx.max_over_ranks(dt, per_arm[0], per_arm[1])
    value = 2 * half * T / dt
    acc = {}
    for a in arms:
        stt, hist = red[a["name"]]
        acc[a["name"]] = dict(parallel.derive_accuracy(stt, half), per_run_avg_pos_err_m=parallel.histogram_summary(hist, 0.0, 10.0))
    ceil_tf = fp64_ceiling(torch)
    ukf_st = red["ukf"][0]
    ukf_tf = ukf_st[9] / (arm_ukf * 1e-3) / 1e12
    rec = {"value": value, "unit": UNIT, "steps": 1, "warmup": W, "ms_per_step": dt * 1e3, "scaling": "strong", "dtype": "f64",
           "config": {"workload": f"{total}-instance mixed Monte-Carlo sweep: {half} EKF-SLAM + {half} UKF-SLAM instances over {cx.world} GPU(s) "
                                  f"(STRONG split, both kinds cut evenly over the ranks), 50-landmark 5x10 grid map, {T} filter steps "
                                  "(BASELINE configs[4])",
                      "instances_total": 2 * half, "instances_this_rank": {a["name"]: a["count"] for a in arms}, "filter_steps": T,
                      "timed_region": "reset + both sweeps + all-reduce (SUM) of the 14-entry statistics vector and of the 1002-bin per-run "
                                      "error histogram of each kind; host wall clock between barrier + synchronize brackets, max over ranks",
                      "l2": "covariance working set of either kind exceeds the 126 MB L2 and is rewritten every step"},
           "warmup_note": f"{W} warm-up sweeps of {warm_T} filter steps each, then 1 timed sweep of {T}",
           "clocks": clk, "gpu_launches": int(launches),
           "arm_ms": {"ekf": arm_ekf, "ukf": arm_ukf},
           "roofline": {"bound": "tensor", "kernel": "ukf step kernels (the UKF half carries the sweep: the EKF half finishes under it)",
                        "achieved": ukf_tf, "peak": ceil_tf, "unit": "TFLOP/s", "frac": ukf_tf / ceil_tf if ceil_tf > 0 else None,
                        "peak_source": FP64_SRC, "traffic": None,
                        "frac_meaning": "nominal UKF flops (SURVEY 8d) of all ranks / UKF arm device time (max over ranks) / (world x ceiling)"
                        if cx.world > 1 else "nominal UKF flops (SURVEY 8d) / UKF arm device time / measured DGEMM ceiling"},
           "e2e": None, "cpu_baseline": None,
           "note": "e2e and cpu_baseline of the two kinds are reported by the top-level (EKF) record and by configs.ukf",
           "accuracy": acc}
    if cx.world > 1 and rec["roofline"]["frac"] is not None:
        rec["roofline"]["frac"] /= cx.world
    for a in arms:
        a["sim"].close(); a["fb"].close()
    del arms
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    cx = Ctx(args)
    K, W = args.steps, args.warmup
    T = args.filter_steps
    which = args.filter
    out = None
    if which in ("all", "ekf"):
        rec = measure_batch(cx, "ekf", args.instances, T, K, W, T, 0 if args.no_e2e else max(1, min(K, args.e2e_sweeps)), T,
                            args.ref_instances_per_core or 16, not args.no_cpu_baseline, True,
                            knobs=([(3, 1)] if args.no_sweep else []) + ([(2, args.cta_threads)] if args.cta_threads else []))
        out = rec
    elif which == "ukf":
        rec = measure_batch(cx, "ukf", args.instances, T, K, W, T, 0 if args.no_e2e else 1, min(T, 200),
                            args.ref_instances_per_core or 8, not args.no_cpu_baseline, True,
                            knobs=[(7, args.ukf_gen)] if args.ukf_gen else [])
        out = rec
    elif which == "large":
        out = measure_large(cx, args.large_landmarks, args.large_discovery, args.large_window, 0 if args.no_e2e else 200,
                            3, not args.no_cpu_baseline)
    elif which == "mixed":
        out = measure_mixed(cx, args.mixed_instances, T, 3, min(T, 20))
    if which == "all" and not args.no_sub:
        subs = {}
        subs["ukf"] = measure_batch(cx, "ukf", args.instances, T, 1, 3, min(T, 30), 0 if args.no_e2e else 1, min(T, 200),
                                    8, not args.no_cpu_baseline, False)
        subs["large"] = measure_large(cx, args.large_landmarks, args.large_discovery, args.large_window, 0 if args.no_e2e else 200,
                                      3, not args.no_cpu_baseline)
        subs["mixed"] = measure_mixed(cx, args.mixed_instances, T, 3, min(T, 20))
        out["configs"] = subs
    if cx.rank == 0:
        head = {"metric": METRIC, "value": out["value"], "unit": UNIT, "n_gpus": cx.world, "steps": out["steps"], "warmup": out["warmup"],
                "ms_per_step": out["ms_per_step"], "higher_is_better": True, "scaling": out["scaling"], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic"}
        for k, v in out.items():
            if k not in head:
                head[k] = v
        print(json.dumps(head), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filter", default="all", choices=["all", "ekf", "ukf", "large", "mixed"],
                    help="all (default): BASELINE configs[1] at the top level + configs.ukf / .large / .mixed as bounded sub-records")
    ap.add_argument("--instances", type=int, default=4096, help="filter instances per GPU (configs[1] / configs[2])")
    ap.add_argument("--filter-steps", type=int, default=1000)
    ap.add_argument("--max-meas", type=int, default=8)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--e2e-sweeps", type=int, default=2)
    ap.add_argument("--ref-instances-per-core", type=int, default=0, help="CPU baseline sample (0 = 16 for EKF, 8 for UKF)")
    ap.add_argument("--no-sweep", action="store_true", help="per-step launches on the value path (to profile ekf_step_kernel)")
    ap.add_argument("--cta-threads", type=int, default=0, help="force the CTA width of the EKF kernels (tuning)")
    ap.add_argument("--ukf-gen", type=int, default=0, help="UKF step generation (slam_tune key 7); 0 = library default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="with --filter all: skip the configs.* sub-records")
    ap.add_argument("--large-landmarks", type=int, default=2000)
    ap.add_argument("--large-discovery", type=int, default=2700, help="untimed map-discovery steps before the large-map window")
    ap.add_argument("--large-window", type=int, default=300, help="timed steady-state filter steps of the large-map record")
    ap.add_argument("--mixed-instances", type=int, default=65536, help="TOTAL instances of the mixed sweep (all ranks)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: NCCL_DEBUG=VERSION (set in this image) makes NCCL print its banner on stdout
    # whatever NCCL_DEBUG_FILE says, so the banner is switched off; INFO / TRACE logs are sent to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
