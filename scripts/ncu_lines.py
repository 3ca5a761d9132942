#!/usr/bin/env python
"""Attribute ncu warp-stall samples / executed instructions of one captured launch to CUDA source lines.

  python scripts/ncu_lines.py gpurun_out/prof_step.ncu-rep ekf_stream 'ekf_stream_kernelILi2' [launch_index] [top]

The SASS page of the report is matched (by instruction order) with `nvdisasm -g` line info of the cubin that
`cuobjdump -xelf` extracts from live_ekf_slam_b200/libslam_filter.so (built with -lineinfo).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, unit, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "live_ekf_slam_b200", "libslam_filter.so")], cwd=tmp,
                   capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(unit + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # instructions of the kernel with their source line
    lines, cur, inside = [], None, False
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = kern in ln
            continue
        if ln.startswith("\t.section") or ln.startswith(".section"):
            inside = False
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    body = [r for r in rows[2:] if len(r) == len(h)]
    si, ie = h.index("# Samples"), h.index("Instructions Executed")
    if len(body) != len(lines):
        print(f"warning: {len(body)} SASS rows in the report vs {len(lines)} in the cubin", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0])
    for r, loc in zip(body, lines):
        agg[loc][0] += int(r[si] or 0)
        agg[loc][1] += int(r[ie] or 0)
    ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    srcs = {}
    print(f"# {os.path.basename(rep)} launch {launch}: {ts} samples, {ti} warp instructions")
    print(f"{'samples':>8s} {'%':>6s} {'inst':>10s} {'%':>6s}  location")
    for loc, (sm, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            path = os.path.join(ROOT, "live_ekf_slam_b200", "csrc", loc[0])
            if path not in srcs and os.path.exists(path):
                srcs[path] = open(path).read().splitlines()
            if path in srcs and loc[1] - 1 < len(srcs[path]):
                text = srcs[path][loc[1] - 1].strip()[:90]
        print(f"{sm:8d} {100 * sm / max(ts, 1):6.1f} {ins:10d} {100 * ins / max(ti, 1):6.1f}  {loc[0] if loc else '?'}:{loc[1] if loc else 0}  {text}")


if __name__ == "__main__":
    main()
