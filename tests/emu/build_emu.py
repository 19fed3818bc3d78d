"""TEST INFRASTRUCTURE: builds tests/_build/libqimcifa_emu.so -- the library's own .cu sources compiled by g++ against
tests/emu/include/cuda_runtime.h, with every kernel running on the host (see that header).  The sources are not edited:
this script writes rewritten copies under tests/_build/emu_src/ in which
  kernel<...><<<grid, block, smem, stream>>>(args);   becomes   emu_launch(grid, block, smem, [=]() { kernel<...>(args); });
  extern __shared__ T name[];                          becomes   T* name = (T*)emu_dyn_smem();
  asm volatile("...");  (the instruction-rate micro-benchmark)  becomes   ;
Usage: python tests/emu/build_emu.py   (also called by tests/test_emu_parity.py when the library is out of date)"""
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qimcifa_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "_build")
SRC_DIR = os.path.join(OUT_DIR, "emu_src")
LIB = os.path.join(OUT_DIR, "libqimcifa_emu.so")

LAUNCH = re.compile(r"^(\s*)(\w+(?:<[^<>]*>)?)<<<(.*?)>>>\((.*)\);(\s*\\?)\s*$")
EXTERN_SHARED = re.compile(r"^(\s*)extern\s+__shared__\s+(\w+)\s+(\w+)\[\];")
ASM = re.compile(r"asm\s+volatile\s*\(.*\);")


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite(text):
    out = []
    for line in text.split("\n"):
        m = LAUNCH.match(line)
        if m:
            ind, kern, cfg, args, tail = m.groups()
            grid, block, smem, _stream = split_top(cfg)
            line = f"{ind}emu_launch((unsigned)({grid}), (unsigned)({block}), (size_t)({smem}), [=]() {{ {kern}({args}); }});{tail}"
        else:
            assert "<<<" not in line, line
            m = EXTERN_SHARED.match(line)
            if m:
                line = f"{m.group(1)}{m.group(2)}* {m.group(3)} = ({m.group(2)}*)emu_dyn_smem();"
            elif "asm volatile" in line:
                line = ASM.sub(";", line)
                assert "asm" not in line, line
        out.append(line)
    return "\n".join(out)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def deps():
    return (sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
            + [os.path.join(ROOT, "include", "qimcifa_cuda.h"), os.path.join(HERE, "include", "cuda_runtime.h"),
               os.path.join(HERE, "emu_runtime.cpp"), os.path.abspath(__file__)])


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps()):
        return LIB
    os.makedirs(SRC_DIR, exist_ok=True)
    cpps = []
    for src in sources():
        dst = os.path.join(SRC_DIR, os.path.basename(src)[:-3] + ".cpp")
        with open(src) as f, open(dst, "w") as g:
            g.write(rewrite(f.read()))
        cpps.append(dst)
    cmd = (["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", os.path.join(HERE, "include"),
            "-I", CSRC, "-I", os.path.join(ROOT, "include")] + cpps + [os.path.join(HERE, "emu_runtime.cpp"), "-o", LIB])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("emulator build failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
