"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
oracle on the same inputs.  Bit-exact: same divisors, same tested-guess counts, same coverage."""
import json
import math
import os
import random

import pytest

from oracle import oracle_c as oc
from oracle import oracle_py as op
from qimcifa_b200 import abi

pytestmark = pytest.mark.gpu

PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23)
KINDS = [abi.KERNEL_EXACT, abi.KERNEL_FILTER, abi.KERNEL_AUTO]


@pytest.fixture(scope="module")
def ctx():
    c = abi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ref_traces(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ref_traces.json")))["cases"]


def _prime_below(x, rnd):
    import sympy

    return sympy.prevprime(x - rnd.randrange(0, 1 << 12))


# ---- tables generated on the device (replaces wheel_gen, reference include/qimcifa.hpp:274-308) ----
@pytest.mark.parametrize("level", [5, 6, 7, 8, 9])
def test_device_tables_enumerate_the_unit_group(ctx, level):
    period, ta, tb = ctx.get_tables(level)
    assert period == math.prod(PRIMES[:level])
    assert len(ta) * len(tb) == math.prod(q - 1 for q in PRIMES[:level])
    if level == 5:
        assert ta == list(op.WHEEL11) and tb == [0]
    if level <= 8:
        res = sorted((a + b) % period for a in ta for b in tb)
        want = [v for v in range(period) if all(v % q for q in PRIMES[:level])]
        assert res == want
    else:
        # level 9: 36.5M residues -- check the CRT structure instead of enumerating
        pa, pb = 30030, 17 * 19 * 23
        assert sorted(a % pa for a in ta) == [v for v in range(pa) if all(v % q for q in PRIMES[:6])]
        assert all(a % pb == 0 for a in ta)
        assert sorted(b % pb for b in tb) == [v for v in range(pb) if all(v % q for q in PRIMES[6:9])]
        assert all(b % pa == 0 for b in tb)
        assert len(set(ta)) == len(ta) and len(set(tb)) == len(tb)


# ---- limb kernels (replaces BigInteger operator%, reference include/qimcifa.hpp:369) ----------------
@pytest.mark.parametrize("nbits,g_limbs", [(64, 1), (64, 2), (96, 1), (96, 2), (96, 3), (128, 1), (128, 2), (128, 3)])
def test_divisible_matches_python_ints(ctx, nbits, g_limbs):
    rnd = random.Random(nbits * 10 + g_limbs)
    gmax_bits = min(32 * g_limbs, nbits - 1)
    for trial in range(4):
        p = rnd.randrange(1 << (gmax_bits - 2), 1 << gmax_bits) | 1
        q = rnd.randrange(1 << (nbits - gmax_bits - 1), 1 << (nbits - gmax_bits)) | 1
        n = p * q
        guesses = [p, q | 1, 1, 3, (1 << gmax_bits) - 1]
        guesses += [p + 2, p - 2 if p > 2 else 1, (p << 1) | 1]
        # limb-boundary values and divisors of N +- 1
        guesses += [(1 << 32) - 1, (1 << 32) + 1, (1 << 64) - 1, (1 << 64) + 1, 0xFFFFFFFF00000001, 0x100000000FFFFFFFF]
        for d in (n + 1, n - 1):
            f = next((s for s in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31) if d % s == 0), None)
            if f:
                guesses.append(f)
        guesses += [rnd.randrange(1, 1 << gmax_bits) | 1 for _ in range(3000)]
        # all odd divisors of a smooth-ish number that fit
        guesses = [g for g in guesses if g & 1 and g < (1 << (32 * g_limbs)) and g > 0]
        got = ctx.divisible(n, guesses, g_limbs)
        want = [n % g == 0 for g in guesses]
        assert got == want
        assert got[0]


def test_divisible_rejects_even_guess(ctx):
    with pytest.raises(abi.QcfError):
        ctx.divisible(15, [4], 1)


# ---- the sweep over index intervals vs the oracle ----------------------------------------------------
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("level", [5, 6, 7, 8, 9])
def test_sweep_indices_matches_oracle(ctx, level, kind):
    rnd = random.Random(1000 + level)
    cases = []
    for nbits in (40, 64, 80, 96, 112, 128):
        half = nbits // 2
        p = _prime_below(1 << half, rnd)
        q = _prime_below(1 << (nbits - half), rnd) if nbits % 2 == 0 else _prime_below(1 << (nbits - half), rnd)
        if p == q:
            q = _prime_below(q, rnd)
        p, q = min(p, q), max(p, q)
        ip = oc.backward(p) - 1
        for width in (1, 999, 250_000):
            lo = max(0, ip - rnd.randrange(0, width))
            cases.append((p * q, lo, lo + width - 1, p))
        # an interval without a divisor
        cases.append((p * q, ip + 1, ip + 70_000, None))
    for n, lo, hi, p in cases:
        cnt, hits = oc.intent_sweep_indices(n, level, lo, hi)
        res = ctx.sweep_indices(n, level, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
        assert res.guesses_tested == cnt, (n, lo, hi)
        assert res.guesses_counted == cnt, (n, lo, hi)
        assert ctx.last_hits() == hits, (n, lo, hi)
        assert res.factor == (hits[0] if hits else None)
        if p is not None:
            assert p in hits


@pytest.mark.parametrize("kind", KINDS)
def test_sweep_indices_edges(ctx, kind):
    n = 1000003 * 1000033
    # empty interval
    r = ctx.sweep_indices(n, 5, 10, 9, flags=kind)
    assert (r.found, r.guesses_tested, r.kernel_launches) == (0, 0, 0)
    # single index, a divisor / not a divisor
    i = oc.backward(1000003) - 1
    for level in (5, 9):
        r = ctx.sweep_indices(n, level, i, i, flags=kind | abi.COUNT_ON_DEVICE)
        assert (r.factor, r.guesses_tested, r.guesses_counted) == (1000003, 1, 1)
        r = ctx.sweep_indices(n, level, i + 1, i + 1, flags=kind | abi.COUNT_ON_DEVICE)
        assert r.found == 0 and r.guesses_tested == r.guesses_counted
    # from index 0: value 1 divides everything (the reference never tests indices 0 and 1)
    r = ctx.sweep_indices(n, 5, 0, 100, flags=kind | abi.COUNT_ON_DEVICE)
    assert r.factor == 1 and r.guesses_tested == 101 == r.guesses_counted
    # several divisors in one interval: 17 * 19 * 23 * 1000003
    m = 17 * 19 * 23 * 1000003
    r = ctx.sweep_indices(m, 5, 2, 2000, flags=kind)
    assert ctx.last_hits() == [d for d in (op.forward(k) for k in range(2, 2001)) if m % d == 0]
    assert r.factor == 17
    # level filter removes multiples of 13: 13*13*prime
    m = 169 * 1000003
    assert ctx.sweep_indices(m, 5, 2, 200, flags=kind).factor == 169
    assert ctx.sweep_indices(m, 6, 2, 200, flags=kind).found == 0
    # ragged: interval strictly inside one wheel period, and spanning a period edge
    for lo, hi in ((5, 7), (470, 490), (479, 480), (480, 480)):
        cnt, hits = oc.intent_sweep_indices(n, 7, lo, hi)
        r = ctx.sweep_indices(n, 7, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
        assert (r.guesses_tested, r.guesses_counted) == (cnt, cnt)


@pytest.mark.parametrize("kind", KINDS)
def test_golden_reference_factors(ctx, ref_traces, kind):
    """Every semiprime the shim-compiled reference factored: same factors through the GPU path with the
    intent split; at nodeCount=1, L=5 also the same batches (coverage)."""
    for c in ref_traces:
        n, level = int(c["n"]), c["level"]
        _, _, node_range, batch_count = abi.node_split(n, 1)
        res = ctx.sweep(n, level, 0, batch_count, batches_per_launch=1, flags=kind)
        assert res.found
        f = res.factor
        assert n % f == 0 and 1 < f < n
        if c["found"]:
            assert {f, n // f} == {int(c["found"][0]), int(c["found"][1])}
        if c["found"] and level == 5 and c["node_count"] == 1:
            # same visiting order: the reference swept (batches_done - 1) whole batches + part of the last
            assert f == int(c["found"][0])
            assert abi.count_guesses(5, res.next_batch_hi + 1, batch_count) <= c["count"] <= res.guesses_tested


@pytest.mark.parametrize("kind", KINDS)
def test_full_64bit_sweep_count_and_factor(ctx, kind):
    """Config 1: 64-bit semiprime of two 32-bit primes, default level 7, the whole range in one call."""
    rnd = random.Random(1001)
    for _ in range(3):
        p, q = _prime_below(1 << 32, rnd), _prime_below((1 << 32) - (1 << 29), rnd)
        n = p * q
        _, _, node_range, batch_count = abi.node_split(n, 1)
        res = ctx.sweep(n, 7, 0, batch_count, batches_per_launch=0, flags=kind | abi.COUNT_ON_DEVICE)
        assert res.found and res.factor in (p, q)
        assert res.guesses_tested == res.guesses_counted == abi.count_guesses(7, res.next_batch_hi, batch_count)


@pytest.mark.parametrize("kind", KINDS)
def test_node_split_union_and_found_in_owner_slice(ctx, kind):
    """Config 3 in miniature: nodeCount=8; the node whose slice holds the factor finds it, the others sweep
    their slices without a hit, and the slices tile the single-node range."""
    rnd = random.Random(1003)
    p = _prime_below(1 << 34, rnd)
    q = _prime_below((1 << 34) + (1 << 33), rnd)
    n = p * q
    _, full_range, r, bc = abi.node_split(n, 8)
    owner = None
    total = 0
    for k in range(8):
        lo, hi = abi.node_batches(bc, r, k)
        res = ctx.sweep(n, 6, lo, hi, flags=kind)
        if res.found:
            assert owner is None
            owner = k
            assert res.factor == p
            ip = oc.backward(p) - 1
            assert lo <= (ip - 2) // abi.BIGGEST_WHEEL < hi
        else:
            assert res.batches_done == hi - lo
            total += res.guesses_tested
    assert owner is not None


def test_found_flag_aborts_sweep(ctx):
    n = (2**48 - 59) * (2**48 - 65)
    _, _, r, bc = abi.node_split(n, 1)
    ctx.set_found(1)
    assert ctx.poll_found() == 1
    res = ctx.sweep(n, 5, bc - 64, bc - 60, batches_per_launch=2)
    assert res.aborted == 1 and res.found == 0 and res.batches_done == 0
    ctx.set_found(0)
    assert ctx.poll_found() == 0
    res = ctx.sweep(n, 5, bc - 64, bc - 62, batches_per_launch=2)
    assert res.aborted == 0 and res.batches_done == 2


@pytest.mark.parametrize("nbits", [96, 112, 128])
def test_large_n_round_trip(ctx, nbits):
    """At BASELINE sizes the oracle cannot sweep whole batches; use size-independent properties:
    a planted divisor inside the swept batches is found, the count equals the closed form, and a
    sweep of batches that hold no divisor finds nothing."""
    rnd = random.Random(nbits)
    half = nbits // 2
    p = _prime_below((1 << half) - (1 << (half - 6)), rnd)
    q = _prime_below(1 << half, rnd)
    n = p * q
    bp = (oc.backward(p) - 3) // abi.BIGGEST_WHEEL
    res = ctx.sweep(n, 5, bp - 1, bp + 2, batches_per_launch=3, flags=abi.COUNT_ON_DEVICE)
    assert res.factor == p
    assert res.guesses_tested == res.guesses_counted == abi.count_guesses(5, bp - 1, bp + 2)
    res = ctx.sweep(n, 6, bp + 1, bp + 3, batches_per_launch=1, flags=abi.COUNT_ON_DEVICE)
    assert res.found == 0 and res.batches_done == 2
    assert res.guesses_tested == res.guesses_counted == abi.count_guesses(6, bp + 1, bp + 3)


# ---- the filter kernel (2-adic quotient along arithmetic progressions) ---------------------------------
def _primes_near(x, count, rnd, spread):
    import sympy

    out = set()
    while len(out) < count:
        out.add(sympy.prevprime(x - rnd.randrange(0, spread)))
    return sorted(out)


@pytest.mark.parametrize("level", [5, 6, 7, 8, 9])
@pytest.mark.parametrize("nbits", [96, 112, 128])
def test_filter_kernel_equals_exact_kernel(ctx, nbits, level):
    """Forced filter kernel vs the exact kernel (itself checked against the oracle) on the same index intervals
    near sqrt(N): same divisors, same device-side counts, equal to the closed form.  Divisors are planted in the
    first period, the last period, the ragged ends and the interior."""
    rnd = random.Random(nbits * 100 + level)
    half = nbits // 2
    period = math.prod(PRIMES[:level])
    width = {5: 3_000_000, 6: 30_000_000, 7: 300_000_000, 8: 10_000_000_000, 9: 1_000_000_000_000}[level]  # indices
    top = (1 << half) - rnd.randrange(1 << (half - 8))
    i_hi = oc.backward(top) - 1
    i_lo = i_hi - width
    v_lo, v_hi = oc.forward(i_lo), oc.forward(i_hi)
    spots = [v_lo + 50_000, v_lo + period + 7, (v_lo + v_hi) // 2, v_hi - period - 9, v_hi - 10]
    if level == 9:  # each exact sweep of this interval takes seconds: two spots (ragged bottom, last period)
        spots = [spots[0], spots[3]]
    used_filter = 0
    for spot in spots:
        import sympy

        p = sympy.prevprime(spot)
        q = sympy.prevprime((1 << (nbits - half)) - rnd.randrange(1 << 20))
        if not (v_lo <= p <= v_hi) or p == q:
            continue
        n = p * q
        re = ctx.sweep_indices(n, level, i_lo, i_hi, flags=abi.KERNEL_EXACT | abi.COUNT_ON_DEVICE)
        hits_e = ctx.last_hits()
        rf = ctx.sweep_indices(n, level, i_lo, i_hi, flags=abi.KERNEL_FILTER | abi.COUNT_ON_DEVICE)
        hits_f = ctx.last_hits()
        assert re.kernel_kind == abi.KERNEL_EXACT
        used_filter += rf.kernel_kind == abi.KERNEL_FILTER
        assert hits_f == hits_e and p in hits_f
        assert rf.guesses_counted == re.guesses_counted == rf.guesses_tested == re.guesses_tested
        assert rf.factor == re.factor
        if level < 9:
            ra = ctx.sweep_indices(n, level, i_lo, i_hi, flags=abi.KERNEL_AUTO | abi.COUNT_ON_DEVICE)
            assert ctx.last_hits() == hits_e and ra.guesses_counted == re.guesses_counted
    assert used_filter >= 2


@pytest.mark.parametrize("level", [5, 6])
def test_filter_kernel_matches_oracle(ctx, level):
    """Forced filter kernel against the CPU oracle directly, on intervals the oracle sweeps in seconds."""
    rnd = random.Random(77 + level)
    for nbits in (96, 128):
        half = nbits // 2
        p = _prime_below((1 << half) - rnd.randrange(1 << (half - 4)), rnd)
        q = _prime_below(1 << (nbits - half), rnd)
        n = p * q
        ip = oc.backward(p) - 1
        width = 800_000 if level == 5 else 8_000_000
        lo = ip - rnd.randrange(1, width // 2)
        hi = lo + width
        cnt, hits = oc.intent_sweep_indices(n, level, lo, hi)
        r = ctx.sweep_indices(n, level, lo, hi, flags=abi.KERNEL_FILTER | abi.COUNT_ON_DEVICE)
        assert r.kernel_kind == abi.KERNEL_FILTER
        assert (r.guesses_tested, r.guesses_counted, ctx.last_hits()) == (cnt, cnt, hits)
        assert p in hits


@pytest.mark.parametrize("nbits,width,force", [(80, 20_000_000, "1,0,0"), (80, 20_000_000, "1,0,4"), (96, 4_500_000, "1,0,0"), (128, 4_500_000, "1,0,0")])
def test_degree1_packed_trips_match_oracle(ctx, nbits, width, force):
    """Degree 1 walks packed 16-bit trips (two guesses per add+min on the high halves of the tracked limb, flagged trips
    replayed at full width; qcf_filter_math.cuh).  Forced degree-1 plans -- weak filters at this size, so thousands of
    replays and survivors, several segments at 80 bits -- against the oracle: same divisors, same counts."""
    rnd = random.Random(1600 + nbits + len(force))
    half = nbits // 2
    for _ in range(2):
        p = _prime_below((1 << half) - rnd.randrange(1 << (half - 4)), rnd)
        q = _prime_below(1 << (nbits - half), rnd)
        n = p * q
        ip = oc.backward(p) - 1
        lo = ip - rnd.randrange(1, width // 2)
        hi = lo + width
        cnt, hits = oc.intent_sweep_indices(n, 5, lo, hi)
        assert p in hits
        with abi.options(plan_force=force):
            r = ctx.sweep_indices(n, 5, lo, hi, flags=abi.KERNEL_FILTER | abi.COUNT_ON_DEVICE)
            pl = ctx.last_plan()
        assert r.kernel_kind == abi.KERNEL_FILTER and pl is not None and pl.degree == 1, (r.kernel_kind, pl and pl.degree)
        assert (r.guesses_tested, r.guesses_counted, ctx.last_hits()) == (cnt, cnt, hits)


def test_filter_kernel_many_divisors(ctx):
    """A number with several divisors inside one interval: every one must survive the filter."""
    rnd = random.Random(5150)
    ps = _primes_near((1 << 40) - (1 << 30), 3, rnd, 1 << 24)
    n = ps[0] * ps[1] * ps[2]  # ~120 bits; the three primes and nothing else divide N in the interval
    lo, hi = oc.backward(ps[0]) - 1 - 1000, oc.backward(ps[2]) - 1 + 1000
    for kind in (abi.KERNEL_EXACT, abi.KERNEL_FILTER, abi.KERNEL_AUTO):
        r = ctx.sweep_indices(n, 5, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
        assert ctx.last_hits() == ps
        assert r.guesses_counted == r.guesses_tested


@pytest.mark.parametrize("kind", [abi.KERNEL_EXACT, abi.KERNEL_FILTER, abi.KERNEL_AUTO])
def test_level_primes_apply_to_filter_survivors(ctx, kind):
    """Superset mode: filter chains may run over a lower wheel level, so a true divisor that is a multiple of one of
    the level's remaining primes survives the filter -- it must still never be reported at that level."""
    import sympy

    rnd = random.Random(99)
    r = sympy.prevprime((1 << 44) - rnd.randrange(1 << 30))
    q = sympy.prevprime((1 << 48) - rnd.randrange(1 << 30))
    for prime, lvl_with, lvl_without in ((13, 5, 6), (17, 6, 7), (19, 7, 8), (23, 8, 9)):
        g = prime * r
        n = g * q
        ig = oc.backward(g) - 1
        width = {5: 2_000_000, 6: 20_000_000, 7: 200_000_000, 8: 5_000_000_000}[lvl_with]
        lo, hi = ig - width, ig + width
        res = ctx.sweep_indices(n, lvl_with, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
        assert ctx.last_hits() == [g] and res.guesses_counted == res.guesses_tested
        for lvl in range(lvl_without, 10):
            res = ctx.sweep_indices(n, lvl, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
            assert res.found == 0 and ctx.last_hits() == []
            assert res.guesses_counted == res.guesses_tested == oc.intent_count(lvl, lo, hi)


@pytest.mark.parametrize("kind", KINDS)
def test_guesses_above_2_pow_64(ctx, kind):
    """128-bit N with sqrt(N) within one batch of 2^64: the padded top batches hold guesses above 2^64 (3 limbs)."""
    p, q = 2**64 - 83, 2**64 - 59  # both prime
    n = p * q
    s, full_range, node_range, bc = abi.node_split(n, 1)
    assert oc.forward(bc * abi.BIGGEST_WHEEL + 1) > 2**64 > p
    res = ctx.sweep(n, 5, bc - 2, bc, batches_per_launch=2, flags=kind | abi.COUNT_ON_DEVICE)
    assert sorted(ctx.last_hits()) == [p, q]  # q = 2^64 - 59 lies in the padding above sqrt(N)
    assert res.factor == p and res.guesses_tested == res.guesses_counted == abi.count_guesses(5, bc - 2, bc)
    # an interval entirely above 2^64
    lo = oc.backward(2**64 + 10**6)
    hi = lo + 3_000_000
    g = next(v for v in (oc.forward(i) for i in range(lo + 1000, lo + 2000)) if v % 13)
    m = g * 1000003
    cnt, hits = oc.intent_sweep_indices(m, 6, lo, hi)
    res = ctx.sweep_indices(m, 6, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
    assert hits == [g] == ctx.last_hits() and res.guesses_counted == cnt == res.guesses_tested


@pytest.mark.parametrize("kind", KINDS)
def test_square_prime_and_tiny_inputs(ctx, kind):
    # perfect square of a prime: the root itself is the last guess of the range
    p = 4294967291
    _, _, _, bc = abi.node_split(p * p, 1)
    assert ctx.sweep(p * p, 5, 0, bc, flags=kind).factor == p
    # a prime N: the whole range holds no divisor; count equals the closed form
    n = 18446744073709551557  # 2^64 - 59
    _, _, _, bc = abi.node_split(n, 1)
    res = ctx.sweep(n, 6, 0, bc, flags=kind | abi.COUNT_ON_DEVICE)
    assert res.found == 0 and res.batches_done == bc and res.guesses_counted == res.guesses_tested == abi.count_guesses(6, 0, bc)
    # tiny N: every guess exceeds N, nothing divides
    for tiny in (35, 221, 17 * 19):
        res = ctx.sweep_indices(tiny, 5, 2, 5000, flags=kind | abi.COUNT_ON_DEVICE)
        want = [d for d in (op.forward(k) for k in range(2, 5001)) if tiny % d == 0]
        assert ctx.last_hits() == want and res.guesses_counted == 4999
    with pytest.raises(abi.QcfError):
        ctx.sweep(1, 5, 0, 1)
    with pytest.raises(abi.QcfError):
        ctx.sweep(35, 4, 0, 1)
    with pytest.raises(abi.QcfError):
        ctx.sweep(35, 11, 0, 1)  # levels 5..10 exist (10 = prime 29, src/qimcifa.cpp:64)


# ---- common-factor mode (QCF_MODE_GCD): the reference built with IS_RSA_SEMIPRIME off ------------------------------
@pytest.mark.parametrize("level", [5, 6, 8])
def test_gcd_mode_matches_oracle_on_short_intervals(ctx, level):
    """Row kernel: every guess takes the per-guess gcd test.  Same guesses reported, same counts as the oracle."""
    import math

    rnd = random.Random(400 + level)
    for trial in range(6):
        small = rnd.choice((29, 31, 37, 41, 1009, 65537))
        n = small * _prime_below(1 << rnd.choice((20, 31, 40)), rnd) * _prime_below(1 << rnd.choice((24, 33, 50)), rnd)
        if trial == 3:
            n *= 8  # even N: the kernels work modulo the odd part
        lo = rnd.randrange(2, 1 << rnd.choice((12, 24, 30)))
        hi = lo + rnd.randrange(20_000, 200_000)
        cnt, hits = oc.intent_sweep_indices(n, level, lo, hi, max_hits=64, gcd_mode=True)
        r = ctx.sweep_indices(n, level, lo, hi, flags=abi.MODE_GCD | abi.COUNT_ON_DEVICE)
        assert (r.guesses_tested, r.guesses_counted) == (cnt, cnt)
        if r.n_hits <= 64:
            assert ctx.last_hits() == hits
        if hits:
            # the divisor reported: gcd of N with the first guess in visiting order (top batch first, ascending inside)
            _, all_hits = oc.intent_sweep_indices(n, level, lo, hi, max_hits=1 << 16, gcd_mode=True)
            first = min(all_hits, key=lambda g: (-((oc.backward(g) - 1 - 2) // oc.BW), g))
            assert r.found and r.factor == math.gcd(first, n)
        else:
            assert not r.found


def test_gcd_mode_chain_kernel_finds_planted_multiples(ctx):
    """Chain kernel (Montgomery product of 256 guesses, one gcd per block): N = p * q * r with large primes, interval of
    2^21 periods near 2^44.  Every guess that is a multiple of p inside the interval must be reported, nothing else."""
    import math

    rnd = random.Random(77)
    p, q, r3 = _prime_below(1 << 24, rnd), _prime_below(1 << 45, rnd), _prime_below(1 << 46, rnd)
    n = p * q * r3  # 115 bits
    i_hi = oc.backward(1 << 44) - 1
    i_lo = i_hi - 480 * (1 << 21) - 12345
    v_lo, v_hi = oc.forward(i_lo), oc.forward(i_hi)
    want = sorted(m * p for m in range(v_lo // p, v_hi // p + 2)
                  if v_lo <= m * p <= v_hi and all((m * p) % s for s in PRIMES[:5]))
    assert 50 < len(want)
    res = ctx.sweep_indices(n, 5, i_lo, i_hi, flags=abi.MODE_GCD | abi.COUNT_ON_DEVICE)
    assert res.guesses_counted == res.guesses_tested == oc.intent_count(5, i_lo, i_hi)
    assert res.found and res.factor == p
    assert res.n_hits >= len(want)  # the record was narrowed to the first hits (more than 64 qualify)
    got = ctx.last_hits()
    assert got and all(math.gcd(g, n) == p for g in got)
    # a number with few common-factor guesses: the record is complete and equals the multiples of the planted prime
    p2 = _prime_below(1 << 28, rnd)
    n2 = p2 * q
    want2 = sorted(m * p2 for m in range(v_lo // p2, v_hi // p2 + 2)
                   if v_lo <= m * p2 <= v_hi and all((m * p2) % s for s in PRIMES[:5]))
    res2 = ctx.sweep_indices(n2, 5, i_lo, i_hi, flags=abi.MODE_GCD)
    assert ctx.last_hits() == want2 and (res2.found == bool(want2))
    if want2:
        first = min(want2, key=lambda g: (-((oc.backward(g) - 1 - 2) // oc.BW), g))
        assert res2.factor == math.gcd(first, n2) == p2
    # the same planted prime in numbers whose top bits are all ones, as 3-limb and as 4-limb numbers: there the Montgomery
    # accumulator cannot stay unreduced (qcf_montmul_acc LAZY, csrc/qcf_gcd.cu) and the kernel subtracts after every product
    import sympy

    for top in (96, 128):
        n3 = p2 * sympy.prevprime((1 << top) // p2)
        assert n3.bit_length() == top and (n3 >> (top - 20)) == (1 << 20) - 1
        res3 = ctx.sweep_indices(n3, 5, i_lo, i_hi, flags=abi.MODE_GCD | abi.COUNT_ON_DEVICE)
        assert ctx.last_hits() == want2 and (res3.found == bool(want2)), top
        assert res3.guesses_counted == res3.guesses_tested


def test_gcd_mode_power_of_two_and_exact_mode_agree_on_semiprimes(ctx):
    n = 1000003 * 1000033
    _, _, _, bc = abi.node_split(n, 1)
    a = ctx.sweep(n, 5, 0, bc, flags=abi.MODE_GCD)
    b = ctx.sweep(n, 5, 0, bc)
    assert a.found and b.found and a.factor == b.factor == 1000003
    z = ctx.sweep_indices(1 << 70, 5, 2, 500000, flags=abi.MODE_GCD)
    assert not z.found and z.guesses_tested == oc.intent_count(5, 2, 500000)


def test_survivor_queue_overflow_is_tested_on_the_spot(ctx, monkeypatch):
    """The filter kernel queues survivors in shared memory and the block tests them at epoch ends; the epoch length
    comes from the EXPECTED survivor rate.  A survivor that finds the queue full anyway is tested on the spot by the
    lane that met it (nothing is dropped, nothing is repeated): forced here with epochs far too long for a weak forced
    filter.  Same divisors, same counts, the filter kernel ran, and the statistic says the path was taken."""
    rnd = random.Random(606)
    p = _prime_below((1 << 40) - rnd.randrange(1 << 30), rnd)
    q = _prime_below(1 << 44, rnd)
    n = p * q
    ip = oc.backward(p) - 1
    lo, hi = ip - 33_000_000, ip + 100_000_000
    want_cnt, want_hits = oc.intent_sweep_indices(n, 5, lo, hi)
    assert want_hits == [p]
    # degree 3 with 2^9-period strides over this interval: a 6-bit filter, ~600 survivors per block and round of chains
    with abi.options(plan_force="3,9,1", plan_drift=0):
        with abi.options(epoch_trips=60000):
            r = ctx.sweep_indices(n, 5, lo, hi, flags=abi.KERNEL_FILTER | abi.COUNT_ON_DEVICE)
            assert (r.guesses_tested, r.guesses_counted, ctx.last_hits(), r.factor) == (want_cnt, want_cnt, [p], p)
            assert r.kernel_kind == abi.KERNEL_FILTER and ctx.last_queue_spills() > 0  # the on-the-spot path ran
        r = ctx.sweep_indices(n, 5, lo, hi, flags=abi.KERNEL_FILTER | abi.COUNT_ON_DEVICE)
        assert (r.guesses_counted, ctx.last_hits()) == (want_cnt, [p]) and r.kernel_kind == abi.KERNEL_FILTER
        assert ctx.last_queue_spills() == 0  # epochs sized from the survivor rate: the queue holds them all


# ---- level 10 (prime 29): listed and prompted by the reference (src/qimcifa.cpp:64), clamped away at :96-98 ----------
@pytest.mark.parametrize("kind", [abi.KERNEL_EXACT, abi.KERNEL_FILTER, abi.KERNEL_AUTO])
def test_level_10_matches_oracle(ctx, kind):
    """Level 10 runs on the level-9 (or lower) tables and drops the multiples of 29: same guesses counted, same
    divisors reported as the oracle's 10-prime wheel; a divisor that is a multiple of 29 is reported at level 9 only."""
    rnd = random.Random(1010 + kind)
    for nbits, width in ((64, 3_000_000), (96, 300_000_000), (128, 30_000_000)):
        half = nbits // 2
        p = _prime_below((1 << half) - rnd.randrange(1 << (half - 4)), rnd)
        q = _prime_below(1 << (nbits - half), rnd)
        n = p * q
        ip = oc.backward(p) - 1
        lo = ip - rnd.randrange(1, width // 2)
        hi = lo + width
        if width <= 30_000_000:
            cnt, hits = oc.intent_sweep_indices(n, 10, lo, hi)
        else:  # the oracle loop would take minutes: closed-form count, the planted divisor
            cnt, hits = oc.intent_count(10, lo, hi), [p]
        r = ctx.sweep_indices(n, 10, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
        assert (r.guesses_tested, r.guesses_counted, ctx.last_hits(), r.factor) == (cnt, cnt, hits, p), (nbits, kind)
        assert cnt < oc.intent_count(9, lo, hi)
    # a divisor that is a multiple of 29
    import sympy

    r29 = sympy.prevprime((1 << 40) - rnd.randrange(1 << 30))
    q = sympy.prevprime((1 << 46) - rnd.randrange(1 << 30))
    n = 29 * r29 * q
    g = 29 * r29
    ig = oc.backward(g) - 1
    lo, hi = ig - 2_000_000, ig + 2_000_000
    at9 = ctx.sweep_indices(n, 9, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
    assert g in ctx.last_hits() and at9.guesses_counted == oc.intent_count(9, lo, hi)
    at10 = ctx.sweep_indices(n, 10, lo, hi, flags=kind | abi.COUNT_ON_DEVICE)
    assert g not in ctx.last_hits() and all(h % 29 for h in ctx.last_hits())
    assert at10.guesses_counted == at10.guesses_tested == oc.intent_count(10, lo, hi)


def test_level_10_whole_batches_and_gcd_mode(ctx):
    n = 1000003 * 1000033
    _, _, _, bc = abi.node_split(n, 1)
    r = ctx.sweep(n, 10, 0, bc, flags=abi.COUNT_ON_DEVICE)
    assert r.found and r.factor == 1000003
    assert r.guesses_counted == r.guesses_tested == abi.count_guesses(10, 0, bc) == oc.intent_count(10, 2, bc * abi.BIGGEST_WHEEL + 1)
    n2 = 31 * 1000003 * 1000033
    lo, hi = 5000, 405000
    cnt, hits = oc.intent_sweep_indices(n2, 10, lo, hi, max_hits=1 << 16, gcd_mode=True)
    g = ctx.sweep_indices(n2, 10, lo, hi, flags=abi.MODE_GCD | abi.COUNT_ON_DEVICE)
    assert (g.guesses_tested, g.guesses_counted) == (cnt, cnt) and g.n_hits == len(hits) and g.found
    with pytest.raises(abi.QcfError):
        ctx.sweep_random(n, 10, 1, 0, 1024)


# ---- exact chain kernel: two-stage loop (low limb first, full test for what passes) vs the one-stage loop ---------
def _stage1_false_positive(g, ln, rnd):
    """N < 2^(32 ln) with N // g < 2^64, g not a divisor, for which stage 1 of the two-stage test passes:
    N = T2 W^2 - M g with T2 = g + j W (tests/test_device_math.py checks the construction on the host build)."""
    w = 1 << 32
    while True:
        j = rnd.randrange(-(g >> 32), w)
        t2 = g + j * w
        if j == 0 or t2 <= 0:
            continue
        lo = max(0, t2 * w * w - (1 << (32 * ln)) + 1)
        m_lo, m_hi = lo // g + 1, min(w * w - 1, (t2 * w * w - 1) // g)
        if m_lo > m_hi:
            continue
        m = rnd.randrange(m_lo, m_hi + 1)
        n = t2 * w * w - m * g
        if 0 < n < (1 << (32 * ln)) and n.bit_length() > 32 * (ln - 1) and n // g < (1 << 64) and n % g and n % 2:
            return n


@pytest.mark.parametrize("level", [5, 7])
@pytest.mark.parametrize("nbits", [96, 128])
def test_two_stage_exact_chain_equals_one_stage(ctx, monkeypatch, nbits, level):
    """xchain_two_stage = 1 (qcf_lowlimb_q2g2 in the loop, full test out of line) against =0 (the full test per guess,
    itself checked against the oracle above) on intervals long enough for the chain form, with a ragged number of
    steps per chain: divisors planted at both ends, in the interior and in the last partial trip; and numbers built so
    that a non-divisor passes stage 1 (the second stage must reject it)."""
    import sympy

    rnd = random.Random(nbits * 10 + level)
    half = nbits // 2
    period = math.prod(PRIMES[:level])
    periods = (1 << 19) + (1 << 18) + (1 << 16) + 12345  # chain steps not a multiple of the trip length
    top = (1 << half) - rnd.randrange(1 << (half - 8))
    i_hi = oc.backward(top) - 1
    i_lo = oc.backward(top - periods * period)
    v_lo, v_hi = oc.forward(i_lo), oc.forward(i_hi)
    spots = [v_lo + 3 * period + 11, (v_lo + v_hi) // 2, v_hi - 2 * period - 5, v_hi - 17,
             v_lo + (periods - 70000) * period, v_lo + ((1 << 19) + (1 << 18)) * period + 4321]
    cases = []
    for spot in spots:
        p = sympy.prevprime(spot)
        q = sympy.prevprime((1 << (nbits - half)) - rnd.randrange(1 << 20))
        if v_lo <= p <= v_hi and p != q:
            cases.append((p * q, p))
    ln = (nbits + 31) // 32
    for _ in range(3):  # stage-1 false positives at wheel guesses inside the interval
        g = sympy.prevprime(rnd.randrange(v_lo, v_hi))
        cases.append((_stage1_false_positive(g, ln, rnd), None))
    for n, p in cases:
        out = {}
        for two in ("0", "1"):
            with abi.options(xchain_two_stage=two):
                r = ctx.sweep_indices(n, level, i_lo, i_hi, flags=abi.KERNEL_EXACT | abi.COUNT_ON_DEVICE)
            out[two] = (ctx.last_hits(), r.found, r.factor, r.guesses_tested, r.guesses_counted)
        assert out["0"] == out["1"], (n, p, out)
        assert out["1"][3] == out["1"][4]
        if p is not None:
            assert p in out["1"][0]
        else:
            assert all(n % h == 0 for h in out["1"][0])
