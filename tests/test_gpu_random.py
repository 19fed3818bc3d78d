"""GPU: random-guess Monte-Carlo mode (no reference code exists for it -- SURVEY.md section 0 -- so the checks are
known answers of the generator, the documented counter -> guess map, determinism, and validity of the factor)."""
import math
import random

import pytest

from oracle import oracle_c as oc
from qimcifa_b200 import abi
from tests.philox_ref import philox4x32_10, random_guess

PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert philox4x32_10([0, 0, 0, 0], (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10([0xFFFFFFFF] * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], (0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.fixture(scope="module")
def ctx():
    c = abi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("level", [5, 7, 9])
@pytest.mark.parametrize("nbits", [64, 96, 128])
def test_device_guesses_follow_the_documented_map(ctx, nbits, level):
    rnd = random.Random(nbits + level)
    n = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
    period, ta, tb = ctx.get_tables(level)
    periods = (math.isqrt(n) + 1) // period
    seed, c0 = rnd.randrange(1 << 64), rnd.randrange(1 << 63)
    got = ctx.random_guesses(n, level, seed, c0, 4096)
    want = [random_guess(c0 + i, seed, periods, period, ta, tb) for i in range(4096)]
    assert got == want
    for g in got[:512]:
        assert g == 1 or (all(g % q for q in PRIMES[:level]) and g < periods * period)


@pytest.mark.gpu
def test_random_mode_finds_the_factor_deterministically(ctx):
    import sympy

    p, q = sympy.prevprime(990000), sympy.nextprime(1010000)  # p lies inside the whole periods below sqrt(N)
    n = p * q
    period, ta, tb = ctx.get_tables(5)
    periods = (math.isqrt(n) + 1) // period
    budget = 1 << 22  # ~20x the 2.1e5 guesses below sqrt(N): the factor is drawn with near certainty
    r1 = ctx.sweep_random(n, 5, 42, 0, budget)
    r2 = ctx.sweep_random(n, 5, 42, 0, budget)
    assert r1.found and r1.factor == p and (r2.found, r2.factor, r2.n_hits) == (1, p, r1.n_hits)
    ones = sum(1 for c in range(0, 1 << 12) if random_guess(c, 42, periods, period, ta, tb) == 1)
    assert r1.guesses_tested <= budget and r1.guesses_tested >= budget - 64 * max(1, ones)
    # the hit count equals the number of counters that map to p (checked on the host twin for a prefix)
    r3 = ctx.sweep_random(n, 5, 42, 0, 1 << 16)
    hits = sum(1 for c in range(1 << 16) if random_guess(c, 42, periods, period, ta, tb) == p)
    assert r3.n_hits == hits
    # another seed gives another stream; disjoint counter blocks partition the work (multi-GPU rule)
    a = ctx.sweep_random(n, 5, 42, 0, 1 << 20)
    b = ctx.sweep_random(n, 5, 42, 1 << 20, 1 << 20)
    c = ctx.sweep_random(n, 5, 42, 0, 1 << 21)
    assert a.n_hits + b.n_hits == c.n_hits and a.guesses_tested + b.guesses_tested == c.guesses_tested


@pytest.mark.gpu
def test_random_mode_128bit_budget(ctx):
    """Config 5 in miniature: fixed guess budget on a 128-bit semiprime; a planted small factor is found only
    if drawn, and any reported factor divides N."""
    n = (2**64 - 59) * (2**64 - 83)
    r = ctx.sweep_random(n, 9, 7, 0, 1 << 24)
    assert r.guesses_tested == 1 << 24 and r.found == 0 and r.kernel_launches == 1 and r.device_seconds > 0
    m = 1000003 * (2**89 - 1)  # 109 bits with a small factor: sqrt(m) ~ 2^54.5, never drawn in 2^20 guesses w.h.p.
    r = ctx.sweep_random(m, 5, 7, 0, 1 << 20)
    assert r.found == 0 or m % r.factor == 0
    with pytest.raises(abi.QcfError):
        ctx.sweep_random(35, 5, 1, 0, 10)  # sqrt(N) below one wheel period


# ---- chain-draw mode (QCF_RANDOM_CHAINS): Philox draws row groups of a random window, the filter kernel walks them ----
def chain_draw_round_trip(ctx, abi_mod, nbits, level, seed, samples=2, force=None):
    """The kernel tests exactly the guesses of the documented map (include/qimcifa_cuda.h), checked end to end:
    the map of counters -> guesses comes from the host-side hook (qcf_random_chain_guesses); a guess g* of that list is made a
    divisor of a number N2 with the same batch count (so the same windows) and, checked, the same map for the sample; the
    sweep over the sample's counters must then report g*, count exactly the non-skipped guesses of the list, report nothing
    for a counter range that does not hold the sample, and split additively over disjoint counter ranges."""
    import sympy

    rnd = random.Random(nbits * 1000 + level * 10 + seed)
    half = nbits // 2
    p1 = sympy.prevprime((1 << half) - rnd.randrange(1 << (half - 4)))
    n1 = p1 * sympy.prevprime(p1 - rnd.randrange(1 << (half - 6)))
    opts = {"plan_force": force} if force else {}
    with abi_mod.options(**opts):
        ctx.sweep_random(n1, level, seed, 0, 1, flags=abi_mod.RANDOM_CHAINS)  # the first sample of block 0 (its first counter is 0)
        pl = ctx.last_plan()
        assert pl is not None
        _, ta, _ = ctx.get_tables(5 if pl.period == 2310 else (6 if pl.period == 30030 else level))
        q = 4 * len(ta) * pl.steps  # guess numbers per row-group sample
        assert q * samples <= 1 << 22, "keep the host-side map of the test small: force a shorter chain"
        block = 7  # any counter block: its window is drawn by Philox
        c0 = block << 34
        dump = ctx.random_chain_guesses(n1, level, seed, c0, q * samples)
        live = [g for g in dump if g]
        assert len(live) > 0.5 * len(dump) and all(all(g % s for s in PRIMES[:level]) for g in live[:2000])
        top = oc.forward(abi_mod.node_split(n1, 1)[3] * abi_mod.BIGGEST_WHEEL + 1)  # the last guess of the top batch (whole batches, as the sweep)
        assert 1 < min(live) and max(live) <= top
        assert len(set(dump[:q]) - {0}) == len([g for g in dump[:q] if g])  # one row group: distinct guesses
        res = ctx.sweep_random(n1, level, seed, c0, q * samples, flags=abi_mod.RANDOM_CHAINS)
        assert res.found == 0 and res.guesses_tested == len(live) and res.kernel_kind == abi_mod.KERNEL_FILTER
        # plant: N2 = g* x q2 with the same batch count as N1, and the same map for these counters
        _, _, _, bc1 = abi_mod.node_split(n1, 1)
        planted = None
        for _ in range(40):
            g = rnd.choice(live[q:] if samples > 1 else live)  # a guess of the LAST sample
            q2 = sympy.prevprime(n1 // g)
            n2 = g * q2
            if abi_mod.node_split(n2, 1)[3] == bc1 and ctx.random_chain_guesses(n2, level, seed, c0, q * samples) == dump:
                planted = (g, n2)
                break
        assert planted, "no planted number kept the map of the sample"
        g, n2 = planted
        r_all = ctx.sweep_random(n2, level, seed, c0, q * samples, flags=abi_mod.RANDOM_CHAINS)
        assert r_all.found == 1 and r_all.factor == g and r_all.guesses_tested == len(live)
        assert r_all.n_hits == dump.count(g)
        if samples > 1:
            # the first sample alone does not hold g*; the samples from the second one on do; counter ranges split additively
            r_a = ctx.sweep_random(n2, level, seed, c0, q, flags=abi_mod.RANDOM_CHAINS)
            r_b = ctx.sweep_random(n2, level, seed, c0 + q, q * (samples - 1), flags=abi_mod.RANDOM_CHAINS)
            assert r_a.found == (1 if g in dump[:q] else 0) and r_b.found == 1 and r_b.factor == g
            assert r_a.guesses_tested + r_b.guesses_tested == r_all.guesses_tested
            assert r_a.n_hits + r_b.n_hits == r_all.n_hits
        # determinism; another seed draws other row groups
        again = ctx.sweep_random(n2, level, seed, c0, q * samples, flags=abi_mod.RANDOM_CHAINS)
        assert (again.found, again.factor, again.n_hits, again.guesses_tested) == (r_all.found, r_all.factor, r_all.n_hits, r_all.guesses_tested)
        assert ctx.random_chain_guesses(n1, level, seed + 1, c0, 64) != dump[:64]
    return pl


@pytest.mark.gpu
@pytest.mark.parametrize("nbits,level,force", [(64, 5, "2,12,0"), (64, 7, "2,12,0"), (96, 5, "2,22,0"), (128, 5, "2,24,0"), (128, 6, "0,23,0")])
def test_chain_draw_mode_round_trip(ctx, nbits, level, force):
    chain_draw_round_trip(ctx, abi, nbits, level, seed=11, samples=2, force=force)


@pytest.mark.gpu
def test_chain_draw_mode_128bit_budget(ctx):
    """Config 5 in miniature: a fixed guess budget on a 128-bit semiprime; the free plan (long chains); nothing is found,
    the count is the budget rounded to whole row-group samples, and levels above the chains' own are counted on the device."""
    n = (2**64 - 59) * (2**64 - 83)
    r = ctx.sweep_random(n, 5, 7, 0, 1 << 34, flags=abi.RANDOM_CHAINS)
    pl = ctx.last_plan()
    q = 4 * (480 if pl.period == 2310 else 5760) * pl.steps
    assert r.found == 0 and r.kernel_launches == 1 and r.guesses_tested == ((1 << 34) // q) * q
    r9 = ctx.sweep_random(n, 9, 7, 0, 1 << 30, flags=abi.RANDOM_CHAINS)
    assert r9.found == 0 and 0 < r9.guesses_tested == r9.guesses_counted < (1 << 30)
    with pytest.raises(abi.QcfError):
        ctx.sweep_random(1000003 * 1000033, 5, 1, 0, 10, flags=abi.RANDOM_CHAINS)  # too small for the filter kernel
