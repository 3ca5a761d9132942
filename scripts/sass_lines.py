#!/usr/bin/env python3
"""Attribute the warp-stall samples of an `ncu --set full` capture to CUDA source lines.

usage: scripts/sass_lines.py <report.ncu-rep> <object.o> <kernel substring> [top N]

ncu's CSV export of the source page lists SASS instructions only; nvdisasm -g gives the line of every SASS instruction of the
same function in the same order, so the two are joined by instruction index.  Prints samples per source line (top N) and per
barrier-delimited region."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
sections, cur, name = collections.OrderedDict(), None, None      # one entry per .text section (template instantiation) that matches
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        name = m.group(1) if kern in m.group(1) else None
        if name:
            sections[name] = []
        continue
    if not name:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        sections[name].append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; data = rows[2:]
iS, iSrc = h.index("# Samples"), h.index("Source")
# the instantiation that was profiled: the section with the same number of instructions
lines = next((v for v in sections.values() if len(v) == len(data)), None)
if lines is None:
    lines = max(sections.values(), key=len) if sections else []
    print(f"warning: no section with {len(data)} instructions among {[len(v) for v in sections.values()]}", file=sys.stderr)
agg = collections.Counter()
tot = 0
for k, r in enumerate(data):
    n = int(r[iS]); tot += n
    agg[lines[k] if k < len(lines) else None] += n
print("total samples", tot)
for key, n in agg.most_common(top):
    print("%6d %5.1f%%  %s" % (n, 100.0 * n / tot, key))
