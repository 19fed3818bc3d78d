"""GPU parity at the launch shape the headline is measured on ("What's weak" #1).

`bench.py` times ONE qcf_sweep call per step: `ctx.sweep(n, level, b_hi - B, b_hi, batches_per_launch=B)` with B in the
hundreds to thousands -- one filter-kernel launch over ~10^9 wheel periods, chains of 2^10..2^24 steps cut into several
segments.  The oracle cannot walk such a launch, and the filter is a NECESSARY-condition test whose failure mode is a
silently dropped divisor, so this campaign plants divisors: for 96/112/128-bit N, wheel levels 5 and 9, B = bench's
default (448 at 96 bits since the steps stay inside the top half of every slice), 896, 1856 and 4096, at the top of the range and at the top of slice 7 of 8 (sqrt(N)/8), it makes the exact call
bench.py makes with the divisor g of N = g * q placed

  * at a uniformly random guess of the launch (>= 200 different N in all), and
  * at the places where the kernel's bookkeeping changes: step 0 and step K-1 of the first and of the last chain, the first
    and the last step of every segment, the padded last trip,

and asserts every time: the divisor is reported (and nothing else), it is the factor the sweep returns, and the guesses
counted on the device equal the closed form (reference include/qimcifa.hpp:369: every guess of a batch is decided).  The
plan of every launch (qcf_last_plan) is recorded and each planted divisor is located in it AFTER the fact, and the last
test fails unless the campaign really exercised degrees 2 and 3, at least two segment counts and every special position.
"""
import random

import pytest

from oracle import oracle_c as oc
from qimcifa_b200 import abi

pytestmark = pytest.mark.gpu

BW = abi.BIGGEST_WHEEL
LEVEL_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23)
SEEN = {"degrees": set(), "segments": set(), "tprimes": set(), "positions": set(), "n_values": set(), "random_hits": 0,
        "outside_main_launch": 0, "base_periods": set(), "counted_on_device": 0}


@pytest.fixture(scope="module")
def ctx():
    c = abi.Context(0)
    yield c
    c.close()


def _bench_batches(nbits):
    """What `python bench.py --bits nbits` (default --steps/--warmup of the driver: 20 + 5) uses as batches per step."""
    import bench

    p, q, n = bench.synthetic_semiprime(nbits, 1002)
    return bench.fixed_batches_per_step(abi, n, p, 20, 5, 1)


def _coprime(g, level):
    return all(g % q for q in LEVEL_PRIMES[:level])


def _interval(b_hi, B):
    return oc.forward((b_hi - B) * BW + 2), oc.forward(b_hi * BW + 1)


def _locate(pl, g):
    """Where guess g sits in plan pl (the largest filter launch of the call): a set of position tags."""
    P, tp, K, U = pl.period, pl.tprime, pl.steps, pl.unroll
    m_lo, m_end = abi.limbs_to_int(pl.m_lo_le), abi.limbs_to_int(pl.m_end_le)
    if not (m_lo * P < g < m_end * P):
        return {"outside"}
    m = g // P - m_lo
    c, k = m & ((1 << tp) - 1), m >> tp
    r = g % P
    tags = {"inside"}
    if k == 0:
        tags.add("step_0")
    if k == K - 1:
        tags.add("step_last")
    if c == 0 and r < 64:
        tags.add("first_chain")
    if c == (1 << tp) - 1 and r > P - 64:
        tags.add("last_chain")
    if K % U and k >= (pl.trips - 1) * U:
        tags.add("padded_last_trip")
    k0 = 0
    for j in range(pl.n_seg):
        k1 = k0 + pl.seg_trips[j] * U
        if k == k0:
            tags.add("segment_first_step")
            if j > 0:
                tags.add("inner_segment_first_step")
        if k == min(k1, K) - 1:
            tags.add("segment_last_step")
            if j < pl.n_seg - 1:
                tags.add("inner_segment_last_step")
        k0 = k1
    return tags


def _unit(P, rnd):
    import math

    while True:
        r = rnd.randrange(1, P)
        if math.gcd(r, P) == 1:
            return r


def _near_coprime(g, level, v_lo, v_hi, step):
    """The guess of the level nearest to g in direction `step` (a multiple of the plan's stride keeps chain and residue)."""
    for _ in range(200):
        if v_lo <= g <= v_hi and _coprime(g, level):
            return g
        g += step
    return None


def _targets(pl, level, v_lo, v_hi, rnd):
    """Guesses at the special positions of plan pl (best effort: the plan of the planted N may differ slightly; the
    positions actually hit are established afterwards by _locate on that launch's own plan)."""
    P, tp, K, U = pl.period, pl.tprime, pl.steps, pl.unroll
    base = abi.limbs_to_int(pl.m_lo_le) * P
    S = P << tp
    units = [r for r in range(1, 64) if _coprime(r, level)]
    first_r, last_r = units[0], P - units[0]
    out = []

    def at(c, r, k):
        g = base + (c + (k << tp)) * P + r
        return g if (v_lo <= g <= v_hi and _coprime(g, level)) else None

    cl = (1 << tp) - 1
    for c, rs in ((0, units), (cl, [P - u for u in units])):
        for k in (0, K - 1):
            for r in rs:  # the first unit whose guess is coprime to the level's extra primes too
                g = at(c, r, k)
                if g:
                    out.append(g)
                    break
    k0 = 0
    for j in range(pl.n_seg):
        k1 = k0 + pl.seg_trips[j] * U
        for k in (k0, min(k1, K) - 1):
            for _ in range(40):
                g = at(rnd.randrange(0, 1 << tp), _unit(P, rnd), k)
                if g:
                    out.append(g)
                    break
        k0 = k1
    if K % U:
        for _ in range(3):
            for _ in range(40):
                g = at(rnd.randrange(0, 1 << tp), _unit(P, rnd), rnd.randrange((pl.trips - 1) * U, K))
                if g:
                    out.append(g)
                    break
    return out


def _plant_and_sweep(ctx, g, q, level, b_hi, B, kind=abi.KERNEL_AUTO, count=False):
    """`count`: also count the guesses on the device (the per-chain tally is as costly as the walk itself when the chains run
    over a lower wheel level than requested, so only some sweeps of every launch shape ask for it)."""
    n = g * q
    res = ctx.sweep(n, level, b_hi - B, b_hi, batches_per_launch=B, flags=kind | (abi.COUNT_ON_DEVICE if count else 0))
    hits = ctx.last_hits()
    want = abi.count_guesses(level, b_hi - B, b_hi)
    assert res.found == 1 and res.factor == g and hits == [g], (n, g, level, b_hi, B, res.found, res.factor, hits)
    assert res.guesses_tested == want and (not count or res.guesses_counted == want), (n, level, b_hi, B, res.guesses_counted, res.guesses_tested, want)
    SEEN["counted_on_device"] += bool(count)
    assert res.batches_done == B and not res.aborted
    pl = ctx.last_plan()
    assert pl is not None and res.kernel_kind == abi.KERNEL_FILTER, "the bench shape must run the filter kernel"
    SEEN["degrees"].add(pl.degree)
    SEEN["segments"].add(pl.n_seg)
    SEEN["tprimes"].add(pl.tprime)
    SEEN["base_periods"].add(pl.period)
    SEEN["n_values"].add(n)
    tags = _locate(pl, g)
    SEEN["positions"] |= tags
    if "outside" in tags:
        SEEN["outside_main_launch"] += 1
    return pl, tags


def _setup(nbits, depth, B, rnd):
    """-> (q, b_hi): the launch [b_hi - B, b_hi) lies at sqrt(N) / depth for every N = g * q with g inside it."""
    import sympy

    half = nbits // 2
    centre = int((1 << half) * 0.93) // depth - rnd.randrange(1 << (half - 10))
    b_hi = (oc.backward(centre) - 2) // BW + 1
    v_lo, v_hi = _interval(b_hi, B)
    # depth 1: q just above the launch (g < sqrt(N) < q, q outside the swept values); depth 8: q = 64 g (g = sqrt(N) / 8)
    q = sympy.nextprime(int(v_hi * (1.02 if depth == 1 else 64.5)))
    assert (v_lo * q).bit_length() == (v_hi * q).bit_length() == nbits, (nbits, depth, (v_lo * q).bit_length())
    return q, b_hi


@pytest.mark.parametrize("depth", [1, 8])
@pytest.mark.parametrize("level", [5, 9])
@pytest.mark.parametrize("nbits", [96, 112, 128])
def test_planted_divisors_at_bench_launch_shape(ctx, nbits, level, depth):
    rnd = random.Random(nbits * 100 + level * 10 + depth)
    for B in sorted({_bench_batches(nbits), 896, 1856, 4096}):
        q, b_hi = _setup(nbits, depth, B, rnd)
        v_lo, v_hi = _interval(b_hi, B)
        p_lo, p_hi = (b_hi - B) * BW + 2, b_hi * BW + 1
        pl = None
        n_random = 12  # x 36 (nbits, level, depth, B) combinations = 432 uniformly placed divisors
        for i in range(n_random):
            g = None
            while g is None or not _coprime(g, level):
                g = oc.forward(rnd.randrange(p_lo, p_hi + 1))  # uniform over the launch's wheel-11 indices
            pl, _ = _plant_and_sweep(ctx, g, q, level, b_hi, B, count=(i == 0))
            SEEN["random_hits"] += 1
        # the very first and the very last guess of the launch's index interval (ragged period ends: exact kernel)
        for g in (_near_coprime(v_lo, level, v_lo, v_hi, 2), _near_coprime(v_hi, level, v_lo, v_hi, -2)):
            assert g is not None
            n = g * q
            res = ctx.sweep(n, level, b_hi - B, b_hi, batches_per_launch=B)
            assert res.found == 1 and res.factor == g and ctx.last_hits() == [g], (n, g)
        for g in _targets(pl, level, v_lo, v_hi, rnd):
            _plant_and_sweep(ctx, g, q, level, b_hi, B)


@pytest.mark.parametrize("force", ["1,0,0", "2,0,1", "2,0,32", "3,0,1", "3,0,32", "3,0,0,2", "2,0,0,1"])
def test_planted_divisors_forced_plan_shapes(ctx, force):
    """The same campaign with the planner restricted to one shape (degree, segment count, padded / whole-trip variant),
    so that every degree and both extremes of the segmentation run at a bench-sized launch whatever the cost model
    prefers today.  Degree 1 needs a weak filter at this size: the kernel selection is forced (>= 6 filter bits)."""
    rnd = random.Random(hash(force) & 0xFFFF)
    nbits, level, B = 96, 5, 1856
    q, b_hi = _setup(nbits, 1, B, rnd)
    v_lo, v_hi = _interval(b_hi, B)
    p_lo, p_hi = (b_hi - B) * BW + 2, b_hi * BW + 1
    kind = abi.KERNEL_FILTER if force.startswith("1") else abi.KERNEL_AUTO
    with abi.options(plan_force=force):
        g0 = oc.forward(rnd.randrange(p_lo, p_hi + 1))
        try:
            pl, _ = _plant_and_sweep(ctx, g0, q, level, b_hi, B, kind, count=True)
        except AssertionError:
            if force.startswith("1"):
                pytest.skip("no degree-1 plan at this launch size")
            raise
        want_d = int(force.split(",")[0])
        if pl.degree != want_d:
            assert want_d == 1, (force, pl.degree)  # only degree 1 may be infeasible here
            pytest.skip("degree 1 is infeasible at this launch size; the planner took degree %d" % pl.degree)
        for _ in range(8):
            g = oc.forward(rnd.randrange(p_lo, p_hi + 1))
            _plant_and_sweep(ctx, g, q, level, b_hi, B, kind)
        for g in _targets(pl, level, v_lo, v_hi, rnd):
            _plant_and_sweep(ctx, g, q, level, b_hi, B, kind)


def test_no_divisor_no_report_at_bench_shape(ctx):
    """The other direction: N prime-like for the launch (both factors far outside it) -- nothing is reported, counts hold."""
    import sympy

    for nbits in (96, 128):
        half = nbits // 2
        p = sympy.prevprime(int((1 << half) * 0.75))
        q = sympy.prevprime(1 << half)
        n = p * q
        _, _, _, bc = abi.node_split(n, 1)
        for level in (5, 9):
            res = ctx.sweep(n, level, bc - 1856, bc, batches_per_launch=1856, flags=abi.COUNT_ON_DEVICE)
            assert res.found == 0 and ctx.last_hits() == [] and res.guesses_counted == res.guesses_tested


def test_campaign_covered_every_shape():
    """Runs last: the campaign above must have exercised what it claims (else a planner change silently shrinks it)."""
    assert len(SEEN["n_values"]) >= 200 and SEEN["random_hits"] >= 200, (len(SEEN["n_values"]), SEEN["random_hits"])
    assert SEEN["counted_on_device"] >= 30, SEEN["counted_on_device"]
    assert {2, 3} <= SEEN["degrees"], SEEN["degrees"]
    assert len(SEEN["segments"]) >= 2, SEEN["segments"]
    need = {"inside", "step_0", "step_last", "first_chain", "last_chain", "segment_first_step", "segment_last_step",
            "inner_segment_first_step", "inner_segment_last_step", "padded_last_trip"}
    assert need <= SEEN["positions"], need - SEEN["positions"]
    print("bench-shape campaign:", {k: (sorted(v) if isinstance(v, set) and k != "n_values" else (len(v) if isinstance(v, set) else v))
                                    for k, v in SEEN.items()})
