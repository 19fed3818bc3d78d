"""CPU-only: pins the oracle (oracle/oracle.cpp + its Python-int twin) against the reference's
embedded constants, the facts established in SURVEY.md section 8c, and traces of the unmodified
reference sources compiled with shims (tests/golden/ref_traces.json)."""
import json
import math
import os
import random

import pytest

from oracle import oracle_c as oc
from oracle import oracle_py as op


@pytest.fixture(scope="module")
def wheel_golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "wheel11.json")))


@pytest.fixture(scope="module")
def ref_traces(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ref_traces.json")))["cases"]


def test_wheel11_matches_reference_table(wheel_golden):
    assert list(op.WHEEL11) == wheel_golden["wheel11"]
    assert oc.wheel11() == wheel_golden["wheel11"]
    assert wheel_golden["BIGGEST_WHEEL"] == op.BIGGEST_WHEEL == oc.BW == 2 * 3 * 5 * 7 * 11 * 13 * 17 * 19 * 23
    assert wheel_golden["SMALLEST_WHEEL"] == op.SMALLEST_WHEEL == 2310
    assert wheel_golden["MIN_RTD_LEVEL"] == op.MIN_RTD_LEVEL == 5


def test_forward_backward_kats():
    # SURVEY.md 8c
    assert [op.forward(i) for i in range(7)] == [1, 13, 17, 19, 23, 29, 31]
    assert op.backward(1) == 1 and op.backward(13) == 2
    assert op.backward(2310) == op.backward(2311) == 481
    rnd = random.Random(7)
    for _ in range(2000):
        p = rnd.randrange(0, 1 << rnd.choice((10, 30, 62)))
        v = op.forward(p)
        assert math.gcd(v, 2310) == 1
        assert oc.forward(p) == v
        assert op.backward(v) - 1 == p == oc.backward(v) - 1
        n = rnd.randrange(1, 1 << rnd.choice((12, 40, 64)))
        assert oc.backward(n) == op.backward(n)
    # forward enumerates exactly the integers coprime to 2310, ascending
    vals = [op.forward(i) for i in range(2000)]
    assert vals == [v for v in range(1, vals[-1] + 1) if math.gcd(v, 2310) == 1]


def test_isqrt_matches_floor_sqrt():
    rnd = random.Random(11)
    for bits in (2, 3, 8, 31, 64, 65, 96, 127, 128):
        for _ in range(50):
            n = rnd.randrange(1 << (bits - 1), 1 << bits)
            assert oc.isqrt(n) == math.isqrt(n)
            if bits <= 96:
                assert op.isqrt(n) == math.isqrt(n)
    for s in (3, 2**32 - 5, 2**64 - 59):
        assert oc.isqrt(s * s) == s and oc.isqrt(s * s - 1) == s - 1


def test_node_split_and_dispenser():
    n = 18446524746962277769
    assert oc.node_split(n, 1) == op.node_split(n, 1)
    assert oc.node_split(n, 8) == op.node_split(n, 8)
    # dispenser with nodeRange=3
    assert op.literal_batches(3, 1, 0) == [2, 1, 0] == oc.literal_batches(3, 1, 0)
    lit = {k: op.literal_batches(3, 8, k) for k in range(8)}
    assert lit == {0: [], 1: [], 2: [], 3: [], 4: [11, 10, 9], 5: [8, 7, 6], 6: [5, 4, 3], 7: [2, 1, 0]}
    assert {k: oc.literal_batches(3, 8, k) for k in range(8)} == lit
    # intent: node k owns the k-th slice from the top, descending
    bw = 1000
    big = 10**9 + 7
    _, _, r, bc = op.node_split(big, 8, bw)
    seen = []
    for k in range(8):
        b = op.intent_node_batches(big, 8, k, bw)
        assert b == list(range(bc - k * r - 1, bc - (k + 1) * r - 1, -1))
        seen += b
    assert sorted(seen) == list(range(bc))


@pytest.mark.parametrize("level", [5, 6, 7, 8, 9])
def test_intent_sweep_cpp_vs_python(level):
    rnd = random.Random(100 + level)
    for _ in range(6):
        lo = rnd.randrange(2, 1 << rnd.choice((8, 20, 40, 60)))
        hi = lo + rnd.randrange(0, 3000)
        # a number with a divisor inside the interval
        g = next(op.intent_guesses(level, lo + (hi - lo) // 2, hi + 5000))
        n = g * 1000003
        cnt, hits = op.intent_sweep_indices(n, level, lo, hi)
        assert oc.intent_sweep_indices(n, level, lo, hi) == (cnt, hits)
        assert oc.intent_count(level, lo, hi) == cnt == op.intent_count(level, lo, hi)
    assert oc.intent_sweep_indices(35, level, 5, 4) == (0, [])  # empty interval


def test_intent_batches_tile_the_range():
    bw = 977
    pieces = [op.batch_index_range(b, bw) for b in range(5)]
    assert pieces[0][0] == 2
    for a, b in zip(pieces, pieces[1:]):
        assert a[1] + 1 == b[0]
    # first tested guesses: 17, 19, 23, 29
    assert list(op.intent_guesses(5, 2, 5)) == [17, 19, 23, 29]


@pytest.mark.parametrize("level", [5, 6, 7])
def test_literal_cpp_vs_python_small_bw(level):
    n = 10007 * 10009
    for bw in (480, 1000, 2311):
        for nc, nid in ((1, 0), (2, 1), (3, 2)):
            tr = []
            f, cnt = op.literal_sweep(n, level, nc, nid, bw, trace=tr)
            cf, ccnt, first, _ = oc.literal_sweep(n, level, nc, nid, bw, trace_cap=len(tr) + 1)
            assert (cf, ccnt, first) == (f, cnt, tr)


def test_literal_first_guesses():
    # L6-literal starts 19,23,29,31; L7-literal 23,29,31
    n = 1000003 * 1000033  # one batch: the sweep starts at index 1
    assert oc.literal_sweep(n, 6, max_guesses=4)[2] == [19, 23, 29, 31]
    assert oc.literal_sweep(n, 7, max_guesses=3)[2] == [23, 29, 31]
    assert oc.literal_sweep(n, 5, max_guesses=4)[2] == [17, 19, 23, 29]


def test_tuner_cardinalities():
    n = (2**50 - 27) * (2**50 - 35)
    assert oc.tuner_cardinalities(n) == op.tuner_cardinalities(n)
    assert oc.tuner_batch_count(n) == op.tuner_batch_count(n)


def test_literal_oracle_matches_reference_traces(ref_traces):
    """The shim-compiled UNMODIFIED reference (single thread) and the oracle's literal mode
    test the same guesses in the same order and report the same factor."""
    for c in ref_traces:
        n = int(c["n"])
        fac, tested, first, h = oc.literal_sweep(n, c["level"], c["node_count"], c["node_id"], trace_cap=64)
        assert tested == c["count"], c
        assert h == c["fnv1a64"], c
        assert [str(x) for x in first] == c["first"], c
        if c["found"] is None:
            assert fac is None
        else:
            assert str(fac) == c["found"][0] and str(n // fac) == c["found"][1]
            assert str(fac) == c["last"]


def test_intent_equals_literal_at_level5_single_node(ref_traces):
    """at nodeCount=1, L=5 the literal reference and the intent semantics
    coincide: same batches, same tested set, same factor."""
    for c in ref_traces:
        if c["level"] != 5 or c["node_count"] != 1:
            continue
        n = int(c["n"])
        _, _, r, bc = oc.node_split(n, 1)
        total, fac = 0, None
        for b in range(bc - 1, -1, -1):
            lo, hi = op.batch_index_range(b)
            cnt, hits = oc.intent_sweep_indices(n, 5, lo, hi)
            if hits:
                fac = hits[0]
                total += oc.intent_count(5, lo, oc.backward(fac) - 1)
                break
            total += cnt
        assert str(fac) == c["found"][0]
        assert total == c["count"]


def test_mt_sweep_port_counts():
    n = 1000003 * 1000033
    fac, tested, sec = oc.mt_sweep(n, 5, 0, 1, threads=2, bw=300000)
    assert fac == 1000003
    assert tested == oc.intent_count(5, 2, 300001)


def test_tuner_cardinality_column_matches_reference_tuner(golden_dir):
    """The `cardinality` column the unmodified reference tuner wrote for a 100-bit semiprime (its time/cost columns
    are noise in this snapshot)."""
    rows = open(os.path.join(golden_dir, "ref_tuner_100bit.ssv")).read().strip().split("\n")
    assert rows[0] == "level, cardinality, batch time (s), cost (s)"
    n = 771819207259660166073186517813
    want = oc.tuner_cardinalities(n)
    assert want == op.tuner_cardinalities(n)
    for r in rows[1:]:
        lvl, card = r.split()[:2]
        assert int(card) == want[int(lvl)]
    assert [r.split()[0] for r in rows[1:]] == ["5", "6", "7", "8", "9"]


def test_common_factor_mode_matches_the_reference_build(golden_dir):
    """The oracle's gcd mode (include/qimcifa.hpp:201-210,373-379) against stdout of the UNMODIFIED reference compiled with
    IS_RSA_SEMIPRIME off (tests/golden/make_golden.py, oracle/_ref/qimcifa_ref_gcd): same reported factor."""
    import math

    cases = json.load(open(os.path.join(golden_dir, "ref_traces.json")))["gcd_cases"]
    swept = 0
    for c in cases:
        n = int(c["n"])
        if c["found"] is None:
            assert "Factors: " in c["stdout_tail"]  # pre-check hit (src/qimcifa.cpp:134-141): no sweep
            continue
        fac, tested, first, _ = oc.literal_sweep(n, c["level"], c["node_count"], c["node_id"], gcd_mode=True)
        assert [str(fac), str(n // fac)] == c["found"], c
        swept += 1
        if c["node_count"] == 1 and c["level"] == 5:
            # intent == literal here: the intent sweep of the top batch reports the same guess first
            _, _, _, bc = oc.node_split(n, 1)
            lo, hi = (bc - 1) * oc.BW + 2, bc * oc.BW + 1
            _, hits = oc.intent_sweep_indices(n, 5, lo, min(hi, lo + 3_000_000), max_hits=4, gcd_mode=True)
            if hits:
                assert math.gcd(hits[0], n) == fac
    assert swept >= 4


def test_common_factor_oracle_against_brute_force():
    import math
    import random

    rnd = random.Random(11)
    for _ in range(20):
        n = rnd.choice((29, 31, 37, 1009)) * rnd.randrange(10**6, 10**7) * 2 ** rnd.randrange(0, 3) + rnd.randrange(0, 2)
        lo = rnd.randrange(2, 5000)
        hi = lo + rnd.randrange(100, 4000)
        cnt, hits = oc.intent_sweep_indices(n, 6, lo, hi, max_hits=4096, gcd_mode=True)
        vals = [v for v in (oc.forward(p) for p in range(lo, hi + 1)) if v % 13]
        assert cnt == len(vals)
        assert hits == [v for v in vals if math.gcd(v, n) != 1]
