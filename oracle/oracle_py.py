"""CPU oracle, Python-int twin (TEST INFRASTRUCTURE ONLY).

A restatement, in exact Python integers, of the reverse-wheel trial-division
sweep of vm6502q/qimcifa.  It is the slow, obviously-correct twin of
``oracle/oracle.cpp`` and is used only by ``tests/`` to pin the C++ oracle on
small cases.  Nothing in the product path (``qimcifa_b200/``) may import it.

Each function cites the reference lines it follows (paths relative to the
reference checkout):

* constants                 include/qimcifa.hpp:86-89
* dispenser                 include/qimcifa.hpp:97-129
* floor sqrt (bisection)    include/qimcifa.hpp:162-186
* wheel11 / forward / backward   include/qimcifa.hpp:224-263
* wheel bitsets             include/qimcifa.hpp:265-327
* test + sweep              include/qimcifa.hpp:364-399
* range / node split        src/qimcifa.cpp:128-158
* tuner cardinality         src/qimcifa_tuner.cpp:132-151

Two modes:

``intent``   what the algorithm means: node k owns the k-th slice from the top,
             batch b owns indices [b*BW+2, (b+1)*BW+1], a guess is tested iff
             it is coprime to the wheel primes of the level.
``literal``  single-thread emulation of what the snapshot really does,
             including the reversed-sentinel bug (D1) and the free-running,
             off-by-one wheel phase (D2).

Parity status: the reference ships no tests or golden vectors.  This twin is
pinned (a) against the reference's embedded constants (tests/golden/wheel11.json,
extracted from include/qimcifa.hpp:224-254) and (b) against traces of the
UNMODIFIED reference sources compiled with compat shims (oracle/_ref, see
oracle/Makefile and tests/golden/ref_traces.json).
"""

BIGGEST_WHEEL = 223092870  # include/qimcifa.hpp:86
SMALLEST_WHEEL = 2310  # include/qimcifa.hpp:88
MIN_RTD_LEVEL = 5  # include/qimcifa.hpp:89
PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29)  # src/qimcifa.cpp:64


def _coprime_to(n, primes):
    return all(n % q for q in primes)


# include/qimcifa.hpp:224-254 -- the 480 integers of [1, 2310] coprime to 2310
# (src/wheel_gen.py:41-43 regenerates them the same way).
WHEEL11 = tuple(i for i in range(1, SMALLEST_WHEEL) if _coprime_to(i, PRIMES[:5]))
assert len(WHEEL11) == 480


def forward(p):
    """include/qimcifa.hpp:256-259: value of 0-based rank p among integers coprime to 2310."""
    return WHEEL11[p % 480] + (p // 480) * 2310


def backward(n):
    """include/qimcifa.hpp:261-263: lower_bound rank in wheel11 + 480*(n/2310) + 1 (1-based)."""
    r = n % 2310
    lb = 0
    while lb < 480 and WHEEL11[lb] < r:
        lb += 1
    return lb + 480 * (n // 2310) + 1


def isqrt(n):
    """include/qimcifa.hpp:162-186: bisection floor square root."""
    start, end, ans = 1, n >> 1, 0
    while True:
        mid = (start + end) >> 1
        sqr = mid * mid
        if sqr == n:
            return mid
        if sqr < n:
            start = mid + 1
            ans = mid
        else:
            end = mid - 1
        if not (start <= end):
            break
    return ans


def level_primes(level):
    """Wheel primes beyond 11 for a level in [5, 9] (src/qimcifa.cpp:64,93-98,152-153)."""
    assert 5 <= level <= 9
    return PRIMES[5:level]


def clamp_level(td_level):
    """src/qimcifa.cpp:93-98 (negative = calibration file, handled by the caller)."""
    if -1 < td_level < MIN_RTD_LEVEL:
        return MIN_RTD_LEVEL
    if td_level > 9:
        return 9
    return td_level


def node_split(n, node_count, bw=BIGGEST_WHEEL):
    """src/qimcifa.cpp:128,149,155-158 -> (fullMaxBase, fullRange, nodeRange, batchCount)."""
    s = isqrt(n)
    full_range = backward(s)
    node_range = (((full_range + node_count - 1) // node_count) + bw - 1) // bw
    return s, full_range, node_range, node_count * node_range


# --------------------------------------------------------------------------
# intent semantics (SURVEY.md Appendix A, S-intent)
# --------------------------------------------------------------------------

def intent_node_batches(n, node_count, node_id, bw=BIGGEST_WHEEL):
    """Batches owned by a node, in visiting (descending) order."""
    _, _, r, bc = node_split(n, node_count, bw)
    hi = bc - node_id * r
    lo = bc - (node_id + 1) * r
    return list(range(hi - 1, lo - 1, -1))


def batch_index_range(b, bw=BIGGEST_WHEEL):
    """include/qimcifa.hpp:388-392 at inc==1: p0=b*BW+1, pre-increment, while p<(b+1)*BW+1."""
    return b * bw + 2, (b + 1) * bw + 1  # inclusive


def intent_guesses(level, p_lo, p_hi):
    """All guesses tested in the inclusive index interval, ascending."""
    qs = level_primes(level)
    for p in range(p_lo, p_hi + 1):
        g = forward(p)
        if _coprime_to(g, qs):
            yield g


def intent_sweep_indices(n, level, p_lo, p_hi):
    """-> (tested_count, [divisors found, ascending])."""
    cnt, hits = 0, []
    for g in intent_guesses(level, p_lo, p_hi):
        cnt += 1
        if n % g == 0:  # include/qimcifa.hpp:369
            hits.append(g)
    return cnt, hits


def intent_sweep_node(n, level, node_count, node_id, bw=BIGGEST_WHEEL):
    """Descending batches; the first batch holding a divisor ends the sweep and the
    smallest divisor in that batch is the one reported (single-thread visiting order).
    -> (factor or None, tested_count, batches_swept)."""
    total, swept = 0, 0
    for b in intent_node_batches(n, node_count, node_id, bw):
        lo, hi = batch_index_range(b, bw)
        cnt, hits = intent_sweep_indices(n, level, lo, hi)
        total += cnt
        swept += 1
        if hits:
            return hits[0], total, swept
    return None, total, swept


def count_coprime_upto(x, primes):
    """#{1 <= v <= x : v coprime to every prime} by inclusion-exclusion (exact)."""
    if x <= 0:
        return 0
    total = 0
    k = len(primes)
    for mask in range(1 << k):
        d, bits = 1, 0
        for i in range(k):
            if mask >> i & 1:
                d *= primes[i]
                bits += 1
        total += (-1) ** bits * (x // d)
    return total


def intent_count(level, p_lo, p_hi):
    """Closed-form tested-guess count of an inclusive index interval."""
    if p_hi < p_lo:
        return 0
    primes = PRIMES[:level]
    return count_coprime_upto(forward(p_hi), primes) - count_coprime_upto(forward(p_lo) - 1, primes)


def tuner_cardinalities(n):
    """src/qimcifa_tuner.cpp:133,137-151: the `cardinality` column, levels 5..9."""
    rng = backward(isqrt(n))
    out = {}
    for i in range(MIN_RTD_LEVEL, 10):
        out[i] = rng
        if i < 9:
            q = PRIMES[i]
            rng = (rng * (q - 1) + q - 1) // q
    return out


def tuner_batch_count(n, bw=BIGGEST_WHEEL):
    """src/qimcifa_tuner.cpp:132."""
    return (isqrt(n) + bw - 1) // bw


# --------------------------------------------------------------------------
# literal semantics (SURVEY.md Appendix A, S-literal) -- single thread
# --------------------------------------------------------------------------

def wheel_inc(primes, limit):
    """include/qimcifa.hpp:274-298 -> (bits as int, size)."""
    primes = list(primes)
    radius = 1
    for i in primes:
        radius *= i
    if limit < radius:
        radius = limit
    prime = primes.pop()
    o = [(i % prime) == 0 for i in range(1, radius + 1) if _coprime_to(i, primes)]
    w = 0
    for i, bit in enumerate(o):
        if bit:
            w |= 1 << i
    return w >> 1, len(o)  # output >>= 1U  (:295)


def wheel_gen(primes, limit):
    """include/qimcifa.hpp:300-308."""
    return [wheel_inc(primes[: i + 1], limit) for i in range(len(primes))]


class LiteralWheels:
    """State of GetWheelIncrement (include/qimcifa.hpp:310-327)."""

    def __init__(self, level, n):
        seqs = wheel_gen(list(PRIMES[:level]), n)
        seqs = seqs[MIN_RTD_LEVEL:]  # src/qimcifa.cpp:153
        self.bits = [w for w, _ in seqs]
        self.size = [s for _, s in seqs]

    def increment(self):
        inc = 0
        while True:
            hit = False
            for i in range(len(self.bits)):
                hit = bool(self.bits[i] & 1)
                self.bits[i] >>= 1
                if hit:
                    self.bits[i] |= 1 << (self.size[i] - 1)
                    break
            inc += 1
            if not hit:
                return inc


class LiteralDispenser:
    """include/qimcifa.hpp:97-129 with src/qimcifa.cpp:156-158."""

    def __init__(self, node_range, node_count, node_id):
        self.batch_number = node_id * node_range
        self.batch_bound = (node_id + 1) * node_range
        self.batch_count = node_count * node_range

    def next(self):
        result = self.batch_count - (self.batch_number + 1)
        if self.batch_number >= self.batch_bound:
            return self.batch_bound
        self.batch_number += 1
        return result

    def finish(self):
        self.batch_number = self.batch_bound


def literal_batches(node_range, node_count, node_id):
    """Batches a single thread visits under the literal control flow (D1)."""
    d = LiteralDispenser(node_range, node_count, node_id)
    out = []
    b = d.next()
    while b < d.batch_bound:  # include/qimcifa.hpp:387 -- compares the REVERSED number
        out.append(b)
        b = d.next()
    return out


def literal_sweep(n, level, node_count=1, node_id=0, bw=BIGGEST_WHEEL, trace=None, max_guesses=None):
    """Single-thread literal sweep (include/qimcifa.hpp:384-399).
    -> (factor or None, tested_count).  `trace` (a list) receives every tested guess."""
    _, _, node_range, _ = node_split(n, node_count, bw)
    wheels = LiteralWheels(level, n)
    disp = LiteralDispenser(node_range, node_count, node_id)
    tested = 0
    b = disp.next()
    while b < disp.batch_bound:
        p = b * bw + 1
        end = (b + 1) * bw + 1
        while p < end:
            p += wheels.increment()
            g = forward(p)
            tested += 1
            if trace is not None:
                trace.append(g)
            if n % g == 0:
                return g, tested
            if max_guesses is not None and tested >= max_guesses:
                return None, tested
        b = disp.next()
    return None, tested
