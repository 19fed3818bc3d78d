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
