#!/usr/bin/env python3
"""Per-source-line table from an ncu --set full report: stall samples (with the dominant stall reasons) and warp
instructions executed, joined with nvdisasm line info like ncu_lines.py.
usage: ncu_phase.py <report.ncu-rep> <kernel name substring, e.g. 'k_quadfit<256, 4096, 0,'> <lib.so> <cubin substring> [top N]"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def load(rep, kern, so, cubsub):
    base_name = re.split(r"[<(]", kern)[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", base_name], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    norm = lambda t: re.sub(r"\((?:int|bool)\)", "", t).replace(" ", "")
    s0 = [i for i in starts if norm(kern) in norm(rows[i][1])][0]
    s1 = min([i for i in starts if i > s0] + [len(rows)])
    rows = rows[s0:s1]
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    sass = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if cubsub in f][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    tmpl = re.findall(r"<(\d+),\s*(\d+),\s*(\d+)", kern)
    pat = base_name
    if tmpl:
        a, b, c = tmpl[0]
        pat = base_name + "ILi" + a + "ELi" + b + "ELb" + c + "E"
    start = [i for i, l in enumerate(lines) if re.search(r"\.text\..*" + pat, l)][0]
    cur = None
    off2line = {}
    for l in lines[start + 1:]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if (re.match(r"\s*\.text\.", l) or l.strip().startswith(".section")) and off2line:
            break
        m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m2:
            off2line[int(m2.group(1), 16)] = cur
    base = int(sass[0][0], 16)
    return hdr, sass, base, off2line


def main():
    rep, kern, so, cubsub = sys.argv[1:5]
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    hdr, sass, base, off2line = load(rep, kern, so, cubsub)
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    tot_s = tot_i = 0
    for r in sass:
        key = off2line.get(int(r[0], 16) - base)
        a = agg.setdefault(key, {"s": 0, "i": 0, "st": {}})
        s = int(r[si]) if r[si].isdigit() else 0
        n = int(r[ii]) if r[ii].isdigit() else 0
        a["s"] += s
        a["i"] += n
        tot_s += s
        tot_i += n
        for ci, name in stall_cols:
            v = int(r[ci]) if r[ci].isdigit() else 0
            if v:
                a["st"][name] = a["st"].get(name, 0) + v
    print(f"{kern}: samples {tot_s}, warp instructions {tot_i}")
    src = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:topn]:
        txt = ""
        if key:
            p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
            if p not in src and os.path.exists(p):
                src[p] = open(p).read().splitlines()
            if p in src and key[1] - 1 < len(src[p]):
                txt = src[p][key[1] - 1].strip()
        st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{n}:{100 * v // max(a['s'], 1)}" for n, v in st)
        print(f"{100.0 * a['s'] / max(tot_s, 1):5.1f}%s {100.0 * a['i'] / max(tot_i, 1):5.1f}%i  {key[1] if key else '?':>4}  [{sts:<34}] {txt[:80]}")


if __name__ == "__main__":
    main()
