"""The plans the library makes down one node slice, per bench step (host only, no GPU): every octave and pass of run_values
(csrc/qcf_api.cu) with the planner's modelled cost.  python tools/slice_plans.py NODE COUNT [BATCHES] [STEPS] [BITS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qimcifa_b200 import abi

node, count = int(sys.argv[1]), int(sys.argv[2])
batches = int(sys.argv[3]) if len(sys.argv) > 3 else 896
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 25
bits = int(sys.argv[5]) if len(sys.argv) > 5 else 96
verbose = os.environ.get("V", "0") == "1"
p, q, n = bench.synthetic_semiprime(bits, 1002)
BW = abi.BIGGEST_WHEEL


def fwd(idx):
    return bench._W11[idx % 480] + (idx // 480) * 2310


def plans(v_lo, v_hi):
    """[(guess share, plan or None)] following run_values' octave / pass loop"""
    out = []
    seg_hi = v_hi
    alive = True
    while True:
        seg_lo = v_lo
        if alive and (seg_hi >> 1) + 1 > v_lo:
            seg_lo = (seg_hi >> 1) + 1
        cur = seg_lo
        planned = False
        for _ in range(8):
            if not alive or cur > seg_hi:
                break
            pl = abi.plan_filter(n, 5, cur, seg_hi)
            if not pl:
                break
            planned = True
            first, last = abi.limbs_to_int(pl.m_lo_le) * pl.period, abi.limbs_to_int(pl.m_end_le) * pl.period
            if cur < first:
                out.append((first - cur, None))
            out.append((last - first, pl))
            cur = last
        if not planned and seg_lo > v_lo:
            alive = False
            continue
        if cur <= seg_hi:
            out.append((seg_hi - cur + 1, None))
        if seg_lo <= v_lo:
            break
        seg_hi = seg_lo - 1
    return out


_, _, node_range, bc = abi.node_split(n, count)
lo, hi = abi.node_batches(bc, node_range, node)
tot_all = 0.0
for k in range(steps):
    b_hi = hi - k * batches
    if b_hi - batches < lo:
        break
    v_lo, v_hi = fwd((b_hi - batches) * BW + 2), fwd(b_hi * BW + 1)
    ps = plans(v_lo, v_hi)
    span = float(v_hi - v_lo + 1)
    cost = sum(w * (pl.cost if pl else 12.0) for w, pl in ps) / span
    tot_all += cost
    print("step %2d b_hi %6d depth %.1f cost %.2f launches %d" % (k, b_hi, (int(n ** 0.5)) / v_hi, cost, len(ps)))
    if verbose:
        for w, pl in ps:
            if pl:
                print("    %5.1f%%  D%d%s tp %2d K %6d J %2d fbits %2d align %2d cost %.2f" % (100 * w / span, pl.degree, "p" if pl.packed else " ", pl.tprime, pl.steps, pl.n_seg, pl.fbits, pl.align, pl.cost))
            else:
                print("    %5.1f%%  exact" % (100 * w / span))
print("mean cost %.3f" % (tot_all / max(1, k + 1)))
