"""Calibration of the filter planner's cost model (plan_filter in qimcifa_b200/csrc/qcf_api.cu) on the GPU.

Forces one plan shape after the other (qcf_set_option("plan_force", "degree,tprime,segments")) for the same 960-batch step at several
depths below sqrt(N), times the step and prints one JSON line per shape with the plan's parameters, so that the
model's constants (trip body per degree, chain setup, per-segment cost, survivor replay) can be fitted off line:
    python tools/plan_cost.py > gpurun_out/plan_cost.jsonl ;  python tools/plan_cost.py --fit gpurun_out/plan_cost.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure():
    import bench
    from qimcifa_b200 import abi

    p, q, n = bench.synthetic_semiprime(96, 1002)
    batches = 896
    per = abi.BIGGEST_WHEEL * 2310 // 480
    with abi.Context(0) as ctx:
        info = ctx.device_info()
        sm, khz = info["sm_count"], info["clock_khz"]
        _, _, r, bc = abi.node_split(n, 1)
        ctx.sweep(n, 5, bc - 8 * batches, bc - 7 * batches, batches_per_launch=batches)  # warm-up
        for div in (1, 8, 32, 64):
            b_hi = bc // div - (3 * batches if div == 1 else 0)
            v_hi = (b_hi * abi.BIGGEST_WHEEL + 1) * 2310 // 480
            v_lo = v_hi - batches * per
            # degree 2 twice: in the packed 16-bit form (where the stride allows it) and in the plain one
            shapes = [(0, 0, 0, 1)] + [(d, tp, j, pk) for d in (1, 2, 3) for tp in range(10, 25) for j in (1, 4, 32) for pk in ((1, 0) if d == 2 else (1,))]
            for d, tp, j, pk in shapes:
                abi.set_option("plan_force", "%d,%d,%d" % (d, tp, j))
                abi.set_option("plan_packed", str(pk))
                pl = abi.plan_filter(n, 5, v_lo, v_hi)
                if pl is None or (d and pl.degree != d) or (tp and pl.tprime != tp) or (d == 2 and pk and not pl.packed):
                    continue
                if pl.steps < 64 and d:
                    continue
                best = None
                for _ in range(2):
                    res = ctx.sweep(n, 5, b_hi - batches, b_hi, batches_per_launch=batches)
                    best = res.device_seconds if best is None else min(best, res.device_seconds)
                surv = sum(pl.seg_trips[i] * pl.seg_bound[i] for i in range(pl.n_seg)) / pl.trips / 2.0 ** 32
                m_cov = (abi.limbs_to_int(pl.m_end_le) - abi.limbs_to_int(pl.m_lo_le)) * 2310 / (v_hi - v_lo)
                cyc = best * sm * 4 * khz * 1e3 / (res.guesses_tested / 32.0)
                print(json.dumps({"div": div, "force": [d, tp, j], "packed": pl.packed, "D": pl.degree, "tp": pl.tprime, "K": pl.steps, "drift": pl.drift, "rho": pl.rho,
                                  "trips": pl.trips, "U": pl.unroll, "J": pl.n_seg, "fbits": pl.fbits, "surv": surv,
                                  "covered": m_cov, "model": pl.cost, "ms": best * 1e3, "launches": res.kernel_launches,
                                  "cycles_per_guess": cyc}), flush=True)
        abi.set_option("plan_force", None)
        abi.set_option("plan_packed", None)


def fit(path):
    """Least squares of  cycles/guess = body[D] * padding + (setup[D] + per_segment[D] * J) / K + survivor * rate
    over the forced shapes; prints FilterCostModel's constants (per degree 1..3)."""
    import numpy as np

    rows = [json.loads(l) for l in open(path) if l.startswith("{")]
    rows = [r for r in rows if r["force"][0] and r["covered"] > 0.98 and r["cycles_per_guess"] < 20]
    A, y = [], []
    for r in rows:
        pad = r["trips"] * r["U"] / r["K"]
        row = []
        for d, pk in ((1, 1), (2, 0), (3, 0), (2, 1)):
            on = float(r["D"] == d and bool(r.get("packed", int(d == 1))) == bool(pk))
            row += [pad * on, on / r["K"], on * r["J"] / r["K"]]
        A.append(row + [r["surv"]])
        y.append(r["cycles_per_guess"])
    A, y = np.array(A), np.array(y)
    used = [i for i in range(A.shape[1]) if np.abs(A[:, i]).sum() > 0]
    coef = np.zeros(A.shape[1])
    coef[used], *_ = np.linalg.lstsq(A[:, used], y, rcond=None)
    names = ["body1", "setup1", "segment1", "body2", "setup2", "segment2", "body3", "setup3", "segment3", "body2p", "setup2p", "segment2p", "survivor"]
    print(", ".join("%s=%.4g" % (n, c) for n, c in zip(names, coef)))
    pred = A @ coef
    err = (pred - y) / y
    print("rows %d, rms rel err %.3f, max %.3f" % (len(y), float(np.sqrt((err ** 2).mean())), float(np.abs(err).max())))
    for r, p_ in zip(rows, pred):
        print("div %2d D%d%s tp%2d K%6d J%2d fbits%2d  meas %.3f  fit %.3f" % (
            r["div"], r["D"], "p" if r.get("packed") else " ", r["tp"], r["K"], r["J"], r["fbits"], r["cycles_per_guess"], p_))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--fit":
        fit(sys.argv[2])
    else:
        measure()
