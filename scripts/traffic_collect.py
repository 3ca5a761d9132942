"""gpurun_out/traffic/{which}.csv (ncu) + {which}_probe.json -> profiles/ncu_traffic.json (read by bench.py's roofline.traffic)"""
import csv, json, os, sys
names = {"step": ["ekf_step_kernel"], "sweep": ["ekf_sweep_kernel"], "gemm": ["lm_gemm"], "ukf": ["ukf_front2_kernel", "ukf_eig3_kernel", "ukf_back3_kernel"]}
out = json.load(open("profiles/ncu_traffic.json")) if os.path.exists("profiles/ncu_traffic.json") else {}     # entries without a new capture are kept
for which, kernels in names.items():
    c, pj = f"gpurun_out/traffic/{which}.csv", f"gpurun_out/traffic/{which}_probe.json"
    if not (os.path.exists(c) and os.path.exists(pj)):
        continue
    probe = json.load(open(pj))
    rows = [r for r in csv.reader(open(c)) if len(r) > 5]
    hi = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hi is None:
        continue
    h = rows[hi]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    for k in kernels:
        mine = [r for r in rows[hi + 1:] if k in r[ki]]
        # every launch of the run is captured; the bracketed step / chunk is the LAST one: its launches are the last `take`
        ids = sorted({r[0] for r in mine}, key=int)
        take = {"gemm": 1, "ukf": 1, "step": 2, "sweep": int(probe.get("launches_in_bracket", 1))}[which]
        mine = [r for r in mine if r[0] in ids[-take:]]
        def val(metric):
            tot = 0.0
            for r in mine:
                if r[mi] == metric:
                    v = float(r[vi].replace(",", ""))
                    u = r[ui].lower()
                    tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
            return tot
        rec = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "dram_bytes_read": val("dram__bytes_read.sum"),
               "dram_bytes_write": val("dram__bytes_write.sum"), "gpu_time_ns_under_ncu": val("gpu__time_duration.sum"),
               "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one launch: {probe['launch']}"}
        rec.update({kk: probe[kk] for kk in ("algorithmic_bytes_same_launch", "moved_model_bytes_same_launch", "mean_n")})
        if which == "ukf":
            rec["note"] = "algorithmic / moved-model bytes are those of the whole three-kernel step"
        out[k] = rec
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
