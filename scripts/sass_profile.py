#!/usr/bin/env python3
"""Attribute an `ncu --set full` source-page CSV to the OUTERMOST source line of the kernel's own translation unit.

usage: scripts/sass_profile.py <source_page.csv> <object.o> <kernel substring> [top N]

`ncu -i rep --page source --csv` lists SASS instructions with their sample / executed-instruction counters;
`nvdisasm -gi` gives, for every SASS instruction of the same function in the same order, the chain of inlined frames.
The two are joined by instruction index.  For every line of the kernel's own file (the outermost frame: libm bodies
such as sincos / atan2 are charged to the line that calls them) the script prints warp-level instructions executed,
stall samples and the dominant stall reason; then the same per barrier-free region is left to the reader."""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
own = os.environ.get("OWN") or os.path.basename(obj).replace(".o", ".cu")
sections, name, frames = collections.OrderedDict(), None, []
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        name = m.group(1) if kern in m.group(1) else None
        if name:
            sections[name] = []
        frames = []
        continue
    if not name:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        frames.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        if frames:
            inner = frames[0]
            outer = next((f for f in (frames if os.environ.get('INNER') else reversed(frames)) if f[0] == own), frames[-1])
            sections[name].append((outer, inner))
            last = (outer, inner)
            frames = []
        else:
            sections[name].append(last)
rows = list(csv.reader(open(src_csv)))
hi = next(k for k, r in enumerate(rows) if "# Samples" in r)
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
iS, iSrc = h.index("# Samples"), h.index("Source")
iX = h.index("Instructions Executed") if "Instructions Executed" in h else None
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
lines = next((v for v in sections.values() if len(v) == len(data)), None)
if lines is None:
    lines = max(sections.values(), key=len) if sections else []
    print(f"warning: no section with {len(data)} instructions among {[len(v) for v in sections.values()]}", file=sys.stderr)


def num(s):
    try:
        return int(float(s.replace(",", "") or 0))
    except ValueError:
        return 0


agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_x = 0
for k, r in enumerate(data):
    key = lines[k][0] if k < len(lines) else None
    n, x = num(r[iS]), (num(r[iX]) if iX is not None else 0)
    tot_s += n; tot_x += x
    a = agg[key]
    a[0] += n; a[1] += x
    for s in stalls:
        a[2][s] += num(r[h.index(s)])
print(f"total samples {tot_s}  warp instructions executed {tot_x}  SASS instructions {len(data)}")
print("%8s %6s %12s %6s  %-28s %s" % ("samples", "%", "warp_inst", "%", "line", "top stalls"))
for key, (n, x, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join("%s %d" % (s.replace("stall_", ""), v) for s, v in st.most_common(3) if v)
    print("%8d %5.1f%% %12d %5.1f%%  %-28s %s" % (n, 100.0 * n / max(tot_s, 1), x, 100.0 * x / max(tot_x, 1), key, tops))
print("\n== by warp instructions executed")
for key, (n, x, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%12d %5.1f%%  samples %5.1f%%  %s" % (x, 100.0 * x / max(tot_x, 1), 100.0 * n / max(tot_s, 1), key))
