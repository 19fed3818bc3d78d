"""The planner's own cost model down a 96-bit range (host only, no GPU): python tools/depth_model.py [BATCHES]
For launches of BATCHES reference batches at depth sqrt(N)/r: the plan chosen and its modelled cost in issue cycles per
guess of a lane (fitted to measurements, profiles/r01b_plan_cost_fit.txt) split into body, chain setup, segments and
survivors -- what a slice far below sqrt(N) (the bottom ranks of a nodeCount split) pays for."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qimcifa_b200 import abi

BODY, SETUP, SEG, SURV = {1: 2.35, 2: 2.73, 3: 4.82}, {1: 270.0, 2: 280.0, 3: 258.0}, {1: 40.0, 2: 46.0, 3: 68.0}, 11500.0
batches = int(sys.argv[1]) if len(sys.argv) > 1 else 960
N = 281200098803711 * 281474976710597
sq = math.isqrt(N)
span = batches * 223092870 * 2310 // 480
print("depth     D  tp      K   J fbits |  body  setup  segs  surv | cost")
for r in (1, 1.5, 2, 4, 8, 16, 32, 64, 128):
    v_hi = int(sq / r)
    v_lo = max(v_hi - span, v_hi // 2 + 1)
    pl = abi.plan_filter(N, 5, v_lo, v_hi)
    if pl is None:
        print("1/%-6s no plan" % r)
        continue
    d, k = pl.degree, pl.steps
    kpad = pl.trips * pl.unroll
    worst = max(pl.seg_bound[: pl.n_seg])
    print("1/%-6s %2d  %2d %6d  %2d  %3d  | %5.2f  %5.2f %5.2f %5.2f | %.3f" % (
        r, d, pl.tprime, k, pl.n_seg, pl.fbits, BODY[d] * kpad / k, SETUP[d] / k, SEG[d] * pl.n_seg / k, SURV * worst / 2**32, pl.cost))
