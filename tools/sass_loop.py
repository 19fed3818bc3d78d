"""Static view of a kernel's hot loop: python tools/sass_loop.py MANGLED [LIB]
Finds the backward branches of the kernel's SASS (cuobjdump), prints the body of the longest loop that contains a multiply
and its opcode histogram.  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "qimcifa_b200", "libqimcifa_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", "-fun", sys.argv[1], lib], capture_output=True, text=True).stdout
ins = []
for ln in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, txt in ins:
    m = re.search(r"BRA\s+(?:\w+,\s*)?(0x[0-9a-f]+)", txt)
    if m and int(m.group(1), 16) <= addr:
        body = [(a, t) for a, t in ins if int(m.group(1), 16) <= a <= addr]
        if any("IMAD" in t or "VIADDMNMX" in t for _, t in body):
            loops.append(body)
which = int(sys.argv[3]) if len(sys.argv) > 3 else None
loops.sort(key=len, reverse=True)
for i, body in enumerate(loops[:6]):
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for _, t in body)
    print(f"# loop {i}: 0x{body[0][0]:x}..0x{body[-1][0]:x}, {len(body)} instructions:", dict(hist.most_common()))
    if which == i:
        for a, t in body:
            print(f"/*{a:04x}*/ {t}")
