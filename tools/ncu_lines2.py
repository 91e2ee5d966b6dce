#!/usr/bin/env python3
"""Like ncu_lines.py, but prints BOTH stall samples and executed warp instructions per source line, and per-file-line-range sums.
usage: ncu_lines2.py <report.ncu-rep> <kernel substring> <lib.so> <cubin name substring> [top N]"""
import csv, re, subprocess, sys, tempfile, os

def main():
    rep, kern, so, cubsub = sys.argv[1:5]
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    base_name = re.split(r"[<(]", kern)[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", base_name], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    norm = lambda t: re.sub(r"\((?:int|bool)\)", "", t).replace(" ", "").replace("b200at::", "")
    sel = [i for i in starts if norm(kern) in norm(rows[i][1])]
    s0 = sel[0]
    s1 = min([i for i in starts if i > s0] + [len(rows)])
    kname = rows[s0][1]
    rows = rows[s0:s1]
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hdr_i]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    sass = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if cubsub in f][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    # mangled name: template ints appear as ILi<n>E / ILb<n>E in order
    nums = re.findall(r"[<,]\s*\(?(?:int|bool)?\)?\s*(\d+)", kname.split("(")[0] if "<" in kname.split("(")[0] else "")
    cands = [i for i, l in enumerate(lines) if re.search(r"\.text\..*" + base_name, l)]
    def ok(l):
        got = re.findall(r"IL[ib](\d+)E", l)
        return got[:len(nums)] == nums
    cs = [i for i in cands if ok(lines[i])] or cands
    start = cs[0]
    cur, per_instr = None, []
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
    base = int(sass[0][0], 16)
    off2line = dict(per_instr)
    agg, tots, toti = {}, 0, 0
    for r in sass:
        s = int(r[si]) if r[si].isdigit() else 0
        n = int(r[ii]) if r[ii].isdigit() else 0
        tots += s; toti += n
        key = off2line.get(int(r[0], 16) - base)
        a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += n
    print(f"{kname[:100]}\n{len(sass)} sass instrs, samples {tots}, warp instructions {toti}")
    src_cache = {}
    for key, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        txt = ""
        if key:
            path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and key[1] - 1 < len(src_cache[path]):
                txt = src_cache[path][key[1] - 1].strip()
        print(f"inst {100.0 * n / max(toti, 1):5.1f}%  stall {100.0 * s / max(tots, 1):5.1f}%  {key}  {txt[:100]}")

if __name__ == "__main__":
    main()
