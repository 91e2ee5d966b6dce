#!/usr/bin/env python3
"""Per-source-region stall samples of one kernel of an ncu report (--set full --import-source on), joining the SASS-level
source page with nvdisasm's line table of the in-tree cubin.
usage: ncu_phase2.py <report.ncu-rep> <kernel index (1-based, as in --kernel-id :::N)> <object.o> <mangled-name substring> [ranges.json]"""
import collections
import csv
import json
import re
import subprocess
import sys
import tempfile
import os

rep, kid, obj, sub = sys.argv[1:5]
ranges = json.load(open(sys.argv[5])) if len(sys.argv) > 5 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if r and r[0].startswith("0x")]
base = int(sass[0][0], 16)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout
fn = None
line = None
off2line = {}
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        fn = m.group(1)
        line = None
        continue
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        line = int(m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+[A-Z@]", l)
    if m and fn and sub in fn:
        off2line[int(m.group(1), 16)] = line
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per_line = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in sass:
    off = int(r[0], 16) - base
    ln = off2line.get(off)
    n = int(r[col["# Samples"]] or 0)
    per_line[ln]["samples"] += n
    per_line[ln]["inst"] += int(r[col["Instructions Executed"]] or 0)
    tot["samples"] += n
    tot["inst"] += int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        v = int(r[col[s]] or 0)
        per_line[ln][s] += v
        tot[s] += v
print("total samples", tot["samples"], "warp-instructions", tot["inst"])
print("stall mix:", ", ".join(f"{s[6:]} {100 * tot[s] / max(1, tot['samples']):.0f}%" for s in sorted(stalls, key=lambda s: -tot[s])[:8]))
if ranges:
    agg = collections.defaultdict(lambda: collections.Counter())
    for ln, c in per_line.items():
        name = "other"
        for a, b, nm in ranges:
            if ln is not None and a <= ln <= b:
                name = nm
                break
        agg[name].update(c)
    for nm, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        top = sorted(stalls, key=lambda s: -c[s])[:3]
        print(f"{100 * c['samples'] / tot['samples']:5.1f}% samples {100 * c['inst'] / tot['inst']:5.1f}% inst  {nm:28s} " + ", ".join(f"{s[6:]} {100 * c[s] / max(1, c['samples']):.0f}%" for s in top))
else:
    for ln, c in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:40]:
        top = sorted(stalls, key=lambda s: -c[s])[:3]
        print(f"{100 * c['samples'] / tot['samples']:5.1f}% line {ln}: inst {c['inst']} " + ", ".join(f"{s[6:]} {100 * c[s] / max(1, c['samples']):.0f}%" for s in top))
