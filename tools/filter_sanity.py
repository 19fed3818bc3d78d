"""Seconds-long GPU sanity check of the default (filter) path: python tools/filter_sanity.py
A 96-bit semiprime whose smaller factor lies 129 batches below sqrt(N) (batch 261914): the top 1984 batches are swept with the
automatic kernel choice (found, counts equal to the closed form, guesses/s), then the batches around the factor with the
forced filter kernel and with the exact kernel (same hits).  No torch, no sympy: starts in a fraction of a second."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qimcifa_b200 import abi

P, Q = 281200098803711, 281474976710597  # prevprime(2^48 - 2^38), prevprime(2^48)
N = P * Q
out = {"n": str(N)}
with abi.Context(0) as ctx:
    _, _, _, bc = abi.node_split(N, 1)
    for _ in range(3):
        r = ctx.sweep(N, 5, bc - 1984, bc, batches_per_launch=1984, flags=abi.KERNEL_AUTO | abi.COUNT_ON_DEVICE)
    out["auto"] = {"found": r.found, "factor_ok": r.factor == P, "kernel_kind": r.kernel_kind, "launches": r.kernel_launches,
                   "counts_equal": r.guesses_tested == r.guesses_counted, "guesses_per_s": r.guesses_tested / r.device_seconds}
    lo = 261914 - 1  # the factor's batch, computed on the host with the oracle's backward()
    hits = {}
    for name, kind in (("filter", abi.KERNEL_FILTER), ("exact", abi.KERNEL_EXACT)):
        r2 = ctx.sweep(N, 5, lo, lo + 4, batches_per_launch=4, flags=kind | abi.COUNT_ON_DEVICE)
        hits[name] = (ctx.last_hits(), r2.guesses_tested, r2.guesses_counted, r2.kernel_kind)
    out["filter_equals_exact"] = hits["filter"][:3] == hits["exact"][:3]
    out["filter_kind"], out["hits"] = hits["filter"][3], [str(h) for h in hits["filter"][0]]
print(json.dumps(out))
