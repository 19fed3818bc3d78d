"""Seconds-long GPU sanity check through every kernel of the library (python tools/all_kernels_sanity.py; written as the workload
for compute-sanitizer, which is closed on this GPU pool -- so it runs plain): a 64-bit semiprime swept
to its factor (filter, exact rows, exact chains, many small launches), the batches around a planted factor of a 96-bit semiprime
with the forced filter kernel (flagged trips, replay by residue class, survivor queue, drain) at degrees 2 and 1 and with the
exact kernel, a degree-3 launch far below sqrt(N), common-factor mode and the chain-draw random mode.  No torch.  Every result
is checked; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qimcifa_b200 import abi

out = {}
with abi.Context(0) as ctx:
    # 64-bit semiprime, the whole sweep down to the factor
    p64, q64 = 4294967291, 4294967279
    n64 = p64 * q64
    _, _, _, bc = abi.node_split(n64, 1)
    r = ctx.sweep(n64, 7, 0, bc, batches_per_launch=1 << 20, flags=abi.COUNT_ON_DEVICE)
    out["sweep64"] = {"found": r.found, "ok": r.factor in (p64, q64), "launches": r.kernel_launches}
    assert r.found and r.factor in (p64, q64)
    # 96-bit semiprime, factor 129 batches below sqrt(N): the batches around it
    P, Q = 281200098803711, 281474976710597
    N = P * Q
    lo = 261914 - 1
    hits = {}
    for name, kind, force in (("filter_d2", abi.KERNEL_FILTER, None), ("filter_d1", abi.KERNEL_FILTER, "1,0,0"), ("filter_d3", abi.KERNEL_FILTER, "3,0,0"),
                              ("exact", abi.KERNEL_EXACT, None)):
        abi.set_option("plan_force", force)
        r2 = ctx.sweep(N, 5, lo, lo + 4, batches_per_launch=4, flags=kind | abi.COUNT_ON_DEVICE)
        hits[name] = (tuple(ctx.last_hits()), r2.guesses_tested, r2.guesses_counted)
        assert r2.found and r2.factor == P and r2.guesses_tested == r2.guesses_counted, (name, r2.found, r2.factor)
    abi.set_option("plan_force", None)
    assert len({h for h in hits.values()}) == 1, hits
    out["planted96"] = {"kernels_agree": True, "hits": [str(h) for h in hits["exact"][0]]}
    # weak filter far below sqrt(N): many survivors through the queue and its drain
    abi.set_option("plan_force", "2,0,4")
    r3 = ctx.sweep(N, 5, 4000, 4008, batches_per_launch=8, flags=abi.KERNEL_FILTER | abi.COUNT_ON_DEVICE)
    abi.set_option("plan_force", None)
    assert not r3.found and r3.guesses_tested == r3.guesses_counted
    out["deep96"] = {"guesses": r3.guesses_tested, "queue_spills": ctx.last_queue_spills() if hasattr(ctx, "last_queue_spills") else None}
    # common-factor mode and random mode
    r4 = ctx.sweep(N, 5, lo, lo + 2, batches_per_launch=2, flags=abi.MODE_GCD | abi.COUNT_ON_DEVICE)
    assert r4.found
    out["gcd"] = {"found": r4.found}
    r5 = ctx.sweep_random(N, 5, 1002, 0, 1 << 26, flags=abi.RANDOM_CHAINS)
    out["random_chains"] = {"guesses": r5.guesses_tested}
print(json.dumps(out))
