"""The planner's own prediction of the 1 -> N GPU curve of `bench.py` (host only, no GPU): for every rank of a nodeCount = N
split, the modelled cost (issue cycles per guess) of each of the bench's steps, from the plan the library would make for that
launch.  python tools/predict_scaling.py [BITS] [STEPS] [WARMUP]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qimcifa_b200 import abi

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 5
p, q, n = bench.synthetic_semiprime(bits, 1002)
B = bench.fixed_batches_per_step(abi, n, p, steps, warm, 1)
BW = abi.BIGGEST_WHEEL


def fwd(idx):
    return bench._W11[idx % 480] + (idx // 480) * 2310


base = None
for world in (1, 2, 4, 8):
    _, _, node_range, bc = abi.node_split(n, world)
    per_rank = []
    for r in range(world):
        lo, hi = abi.node_batches(bc, node_range, r)
        tot = 0.0
        for k in range(warm, warm + steps):
            b_hi = hi - k * B
            v_lo, v_hi = fwd((b_hi - B) * BW + 2), fwd(b_hi * BW + 1)
            pl = abi.plan_filter(n, 5, max(v_lo, v_hi // 2 + 1), v_hi)
            tot += pl.cost if pl else 12.0
        per_rank.append(tot / steps)
    t = max(per_rank)
    base = base or t
    print("N=%d  B=%d  cost/guess by rank: %s  -> predicted efficiency %.3f" % (world, B, " ".join("%.2f" % c for c in per_rank), base / t))
