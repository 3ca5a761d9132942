#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/ (run on the CPU box).

  python scripts/ncu_summary.py full   gpurun_out/prof_ekf.ncu-rep  profiles/r01_ekf_step_full.txt
  python scripts/ncu_summary.py launch gpurun_out/launches_ekf.csv  profiles/r01_ekf_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source: {rep}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:75s} {r[i]:>18s} {units[i]}\n")
        # stall breakdown of the first captured launch (source page)
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        hi = next((k for k, r in enumerate(srows) if "# Samples" in r), None)
        if hi is not None:
            h = srows[hi]
            body = [r for r in srows[hi + 1:] if len(r) == len(h) and r != h and (r[h.index("# Samples")] or "0").isdigit()]
            stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            tot = {s: sum(int(r[h.index(s)] or 0) for r in body) for s in stalls}
            n = sum(int(r[h.index("# Samples")] or 0) for r in body)
            f.write(f"\n== warp-stall samples of the first launch ({n} samples)\n")
            for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
                f.write(f"{s:28s} {v:8d} {100.0 * v / max(n, 1):6.1f} %\n")
            f.write("\n== hottest SASS instructions (samples, instruction, dominant stall)\n")
            for r in sorted(body, key=lambda r: -int(r[h.index('# Samples')] or 0))[:16]:
                best = max(stalls, key=lambda s: int(r[h.index(s)] or 0))
                f.write(f"{r[h.index('# Samples')]:>6s}  {r[h.index('Source')][:80]:80s} {best}\n")


def launch(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[mi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault((r[ki].split("(")[0], r[gi], r[bi]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised; compare SHARES); source: {path}\n")
        f.write(f"{'kernel':60s} {'grid':>14s} {'block':>12s} {'launches':>8s} {'mean_us':>10s} {'share':>7s}\n")
        for (k, g, b), v in agg.items():
            f.write(f"{k[:60]:60s} {g:>14s} {b:>12s} {len(v):8d} {sum(v) / len(v) / 1e3:10.2f} {100 * sum(v) / tot:6.1f}%\n")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
