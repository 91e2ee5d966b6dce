#!/usr/bin/env python3
"""Per-source-line view of one kernel from an ncu report captured with --import-source on:
   ncu_src.py <report.ncu-rep> <kernel regex> [top N]   -> lines sorted by stall samples, with executed warp instructions."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1:3]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = ""
agg = defaultdict(lambda: [0, 0, "", defaultdict(int)])
hdr = None
kname = None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if kname is not None and r[1] != kname:
            break  # first matching kernel only
        kname = r[1]
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
        s, n = int(r[si] or 0), int(r[ii] or 0)
    except ValueError:
        continue
    a = agg[(fname, ln)]
    a[0] += s
    a[1] += n
    a[2] = r[1].strip()
    for i, nm in stall_cols:
        try:
            a[3][nm] += int(r[i] or 0)
        except ValueError:
            pass
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"# {kname}\n# samples {tot_s}, warp instructions {tot_i}")
mix = defaultdict(int)
for a in agg.values():
    for k, v in a[3].items():
        mix[k] += v
print("# stall mix: " + ", ".join(f"{k} {100 * v / tot_s:.0f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{k} {100 * v / max(a[0], 1):.0f}%" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * a[0] / tot_s:5.1f}% smp {100 * a[1] / tot_i:5.1f}% inst  {f}:{ln:<4d} {a[2][:90]}   [{top}]")
