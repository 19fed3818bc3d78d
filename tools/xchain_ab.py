"""A/B of the exact chain kernel's two loops near sqrt(N): python tools/xchain_ab.py [BITS] [BATCHES] [LEVEL]
option xchain_two_stage = 0: full reduction per guess; = 1: low limb first, full test out of line (qcf_exact_chain.cu).
No torch import (the first one on a fresh box takes up to a minute)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sympy

from qimcifa_b200 import abi

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 96
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 256
level = int(sys.argv[3]) if len(sys.argv) > 3 else 5
half = bits // 2
p = sympy.prevprime((1 << half) - (1 << (half - 3)))  # a factor far below the swept batches: nothing is found
q = sympy.prevprime(1 << (bits - half))
n = p * q
out = {"bits": bits, "batches": batches, "level": level, "n": str(n)}
with abi.Context(0) as ctx:
    _, _, _, bc = abi.node_split(n, 1)
    hi = bc - 8
    for two in ("0", "1", "0", "1"):
        abi.set_option("xchain_two_stage", two)
        best = None
        for _ in range(4):
            res = ctx.sweep(n, level, hi - batches, hi, batches_per_launch=batches, flags=abi.KERNEL_EXACT | abi.COUNT_ON_DEVICE)
            assert res.found == 0 and res.guesses_tested == res.guesses_counted
            rate = res.guesses_tested / res.device_seconds
            best = rate if best is None else max(best, rate)
        out.setdefault("two_stage_" + two, []).append(best)
        print("xchain_two_stage=%s  %.3e guesses/s  (%d launches, %.2f ms)" % (two, best, res.kernel_launches, res.device_seconds * 1e3), flush=True)
print(json.dumps(out))
