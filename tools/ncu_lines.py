#!/usr/bin/env python3
"""Join ncu's SASS-level sampling (--page source --csv) with nvdisasm line info: per-source-line stall samples.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> <cubin name substring> [top N]"""
import csv
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, kern, so, cubsub = sys.argv[1:5]
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    base_name = re.split(r"[<(]", kern)[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", base_name], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several kernels (template instances) may be concatenated: pick the section whose "Kernel Name" contains `kern`
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    norm = lambda t: re.sub(r"\((?:int|bool)\)", "", t).replace(" ", "")
    sel = [i for i in starts if norm(kern) in norm(rows[i][1])]
    s0 = sel[0]
    s1 = min([i for i in starts if i > s0] + [len(rows)])
    rows = rows[s0:s1]
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hdr_i]
    si = hdr.index("# Samples")
    sass = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
    mangled_hint = kern
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if cubsub in f][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    # locate the function
    lines = dis.splitlines()
    tmpl = re.findall(r"<(\d+),\s*(\d+),\s*(\d+)", kern)
    pat = base_name
    if tmpl:
        a, b, c = tmpl[0]
        pat = base_name + "ILi" + a + "ELi" + b + "ELb" + c + "E"
    start = [i for i, l in enumerate(lines) if re.search(r"\.text\..*" + pat, l)]
    start = start[0]
    cur = None
    per_instr = []
    for l in lines[start + 1:]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*\.text\.", l) or l.strip().startswith(".section"):
            if per_instr:
                break
        m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m2:
            per_instr.append((int(m2.group(1), 16), cur))
    # join on the instruction offset within the function (ncu prints absolute addresses)
    base = int(sass[0][0], 16)
    off2line = {o: c for o, c in per_instr}
    agg = {}
    tot = 0
    missing = 0
    for r in sass:
        s = int(r[si]) if r[si].isdigit() else 0
        tot += s
        off = int(r[0], 16) - base
        key = off2line.get(off)
        if key is None:
            missing += s
        agg[key] = agg.get(key, 0) + s
    print("samples without line info:", missing)
    print(f"{kern}: {len(sass)} sass instrs, {len(per_instr)} disasm instrs, total samples {tot}")
    src_cache = {}
    for key, s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
        txt = ""
        if key:
            path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and key[1] - 1 < len(src_cache[path]):
                txt = src_cache[path][key[1] - 1].strip()
        print(f"{100.0 * s / max(tot, 1):5.1f}%  {key}  {txt[:110]}")


if __name__ == "__main__":
    main()
