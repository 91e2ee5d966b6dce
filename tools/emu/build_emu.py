#!/usr/bin/env python3
"""Builds tools/emu/_build/libb200apriltags_emu.so: the detector's own .cu sources, translated textually
(kernel<<<...>>>(...) -> emu::launch, extern __shared__ -> emulator buffer) and compiled with g++ against the SIMT
emulator (emu_cuda.h / emu_runtime.cpp).  TEST INFRASTRUCTURE: used by tests/test_emu_parity.py and by developers
without a GPU; never loaded by the package, never timed."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "isaac_ros_apriltag_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libb200apriltags_emu.so")
CUDA_INC = "/usr/local/cuda/include"

EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+((?:__align__\(\d+\)\s+)?)([A-Za-z_][A-Za-z_0-9 ]*?)\s+([A-Za-z_][A-Za-z_0-9]*)\s*\[\s*\]\s*;")
NAME_BEFORE = re.compile(r"([A-Za-z_][A-Za-z_0-9]*(?:<[^<>;(){}]*>)?)\s*$")


def translate(text):
    """NAME<<<CFG>>>(ARGS) -> emu::launch("NAME", CFG, [=]() { NAME(ARGS); })   (balanced-parenthesis scan, newlines kept; the
    lambda copies what the argument expressions refer to, because in the asynchronous mode the launch runs later)."""
    out, pos = [], 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            out.append(text[pos:])
            break
        m = NAME_BEFORE.search(text, pos, i)
        j = text.find(">>>", i)
        assert m and j > 0, "cannot parse kernel launch near: " + text[i - 40:i + 40]
        k = j + 3
        while text[k].isspace():
            k += 1
        assert text[k] == "(", "kernel launch without argument list: " + text[i - 40:i + 40]
        depth, e = 0, k
        while True:
            c = text[e]
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
                if depth == 0:
                    break
            e += 1
        out.append(text[pos:m.start(1)])
        out.append(f"emu::launch(\"{m.group(1)}\", {text[i + 3:j]}, [=]() {{ {m.group(1)}({text[k + 1:e]}); }})")
        pos = e + 1
    text = "".join(out)

    def ext(m):
        ty, name = m.group(2).strip(), m.group(3)
        return f"{ty} *{name} = reinterpret_cast<{ty} *>(emu::dyn_smem());"

    return EXTERN_SHARED.sub(ext, text)


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, opt="-O1", verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h", ".inc"))]
    deps += [os.path.join(HERE, f) for f in ("emu_cuda.h", "emu_runtime.cpp", "build_emu.py")]
    deps.append(os.path.join(ROOT, "include", "b200_apriltags.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    flags = ["-std=c++17", opt, "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-I", CSRC, "-I",
             os.path.join(ROOT, "include"), "-I", CUDA_INC, "-I", HERE]
    objs, procs = [], []
    for f in sources():
        src = os.path.join(CSRC, f)
        gen = os.path.join(OUT_DIR, f[:-3] + ".emu.cpp")
        with open(src) as fh:
            body = translate(fh.read())
        with open(gen, "w") as fh:
            fh.write(f'#line 1 "{src}"\n' + body)
        obj = gen[:-4] + ".o"
        objs.append(obj)
        cmd = ["g++"] + flags + ["-include", os.path.join(HERE, "emu_cuda.h"), "-c", gen, "-o", obj]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    rt_obj = os.path.join(OUT_DIR, "emu_runtime.o")
    procs.append(("emu_runtime.cpp", subprocess.Popen(["g++"] + flags + ["-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", rt_obj],
                                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bad = False
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            bad = True
            sys.stderr.write(f"--- {name} ---\n{out[:6000]}\n")
        elif verbose and out.strip():
            sys.stderr.write(out)
    if bad:
        raise RuntimeError("emulator build failed")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + [rt_obj, "-Wl,--no-undefined", "-Wl,-Bsymbolic", "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
