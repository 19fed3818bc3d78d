"""GPU: randomized stress of every launch path (row-form exact, chain-form exact, filter at degrees 1-3, superset mode,
octave planning, ragged ends) against the CPU oracle: same divisors, same device-side counts, on random N of 34-128
bits, random levels, random index intervals at random depths below sqrt(N)."""
import random

import pytest

from oracle import oracle_c as oc
from qimcifa_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = abi.Context(0)
    yield c
    c.close()


def _odd_cofactor(rnd, bits):
    return rnd.randrange(1 << (bits - 1), 1 << bits) | 1 if bits > 1 else 1


TOTAL_USED = {1: 0, 2: 0}


@pytest.mark.parametrize("seed", range(16))
def test_random_intervals_match_oracle(ctx, seed):
    rnd = random.Random(9000 + seed)
    used = {abi.KERNEL_EXACT: 0, abi.KERNEL_FILTER: 0}
    for trial in range(14):
        level = rnd.randrange(5, 10)
        nbits = rnd.choice((34, 48, 64, 72, 96, 100, 112, 128))
        gbits = rnd.randrange(max(6, nbits // 2 - rnd.choice((0, 0, 1, 3, 8))), nbits // 2 + 1)
        width = rnd.choice((1, 37, 5_000, 400_000, 3_000_000, 12_000_000))
        top = rnd.randrange(1 << (gbits - 1), 1 << gbits)
        i_hi = oc.backward(top) - 1
        i_lo = max(0, i_hi - width)
        # plant a divisor of the level (or none): g * odd cofactor of the complementary size
        n = None
        if rnd.random() < 0.8:
            cands = [oc.forward(i) for i in (rnd.randrange(i_lo, i_hi + 1) for _ in range(40))]
            cands = [g for g in cands if g > 1 and all(g % q for q in (13, 17, 19, 23)[: level - 5])]
            if cands:
                g = cands[0]
                n = g * _odd_cofactor(rnd, max(2, nbits - g.bit_length()))
        if n is None:
            n = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
        if n.bit_length() > 128:
            n >>= n.bit_length() - 128
            n |= 1
        cnt, hits = oc.intent_sweep_indices(n, level, i_lo, i_hi)
        for kind in (abi.KERNEL_AUTO, abi.KERNEL_FILTER, abi.KERNEL_EXACT):
            res = ctx.sweep_indices(n, level, i_lo, i_hi, flags=kind | abi.COUNT_ON_DEVICE)
            got = ctx.last_hits()
            assert got == hits[: abi.MAX_HITS], (seed, trial, n, level, i_lo, i_hi, kind)
            assert res.guesses_tested == cnt == res.guesses_counted, (seed, trial, n, level, i_lo, i_hi, kind)
            if res.kernel_kind in used:
                used[res.kernel_kind] += 1
    assert used[abi.KERNEL_EXACT] > 0
    for k in used:
        TOTAL_USED[k] += used[k]


def test_fuzz_exercised_the_filter_kernel():
    assert TOTAL_USED[abi.KERNEL_FILTER] >= 40, TOTAL_USED


def test_whole_range_sweeps_of_small_numbers(ctx):
    """Octave planning from the very bottom: whole single-node range of 40-56-bit semiprimes, every level."""
    rnd = random.Random(4242)
    import sympy

    for bits in (40, 48, 56):
        p = sympy.prevprime(rnd.randrange(1 << (bits // 2 - 1), 1 << (bits // 2)))
        q = sympy.nextprime(rnd.randrange(1 << (bits // 2 - 1), 1 << (bits // 2)))
        n = p * q
        _, _, _, bc = abi.node_split(n, 1)
        for level in (5, 7, 9):
            for kind in (abi.KERNEL_AUTO, abi.KERNEL_FILTER):
                res = ctx.sweep(n, level, 0, bc, batches_per_launch=bc, flags=kind | abi.COUNT_ON_DEVICE)
                assert res.factor == min(p, q)
                assert res.guesses_tested == res.guesses_counted == abi.count_guesses(level, 0, bc)
