#!/usr/bin/env python3
"""Per-source-line view of one kernel from an ncu report captured with --import-source on:
   ncu_src.py <report.ncu-rep> <kernel regex> [top N] [substring the function name must contain]
   -> lines sorted by stall samples, with their share of the executed warp instructions and the top stall reasons."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1:3]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
must = sys.argv[4] if len(sys.argv) > 4 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}
hdr, fname, func, take, done = None, "", None, False, False
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Kernel Name"):
        if take and func is not None and r[1] != func:
            done = True
        if not done:
            take = must.replace(" ", "") in r[1].replace(" ", "")
            if take:
                func = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not take or done or len(r) < len(hdr) or not r[0].isdigit():
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip(), defaultdict(int)])
    num = lambda t: int(t) if t.isdigit() else 0
    a[0] += num(r[si])
    a[1] += num(r[ii])
    for i, nm in stall_cols:
        a[3][nm] += num(r[i])
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"# {func}\n# samples {tot_s}, warp instructions {tot_i}")
mix = defaultdict(int)
for a in agg.values():
    for k, v in a[3].items():
        mix[k] += v
print("# stall mix: " + ", ".join(f"{k} {100 * v / tot_s:.0f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{k} {100 * v / max(a[0], 1):.0f}%" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * a[0] / tot_s:5.1f}% smp {100 * a[1] / tot_i:5.1f}% inst  {f}:{ln:<4d} {a[2][:90]}   [{top}]")
