"""CPU-only: the arithmetic of the filter kernel (qimcifa_b200/csrc/qcf_filter.cu), restated in exact Python integers,
run on the launch parameters the REAL planner produces (plan_filter in qcf_api.cu, through the qcf_plan_filter hook --
host code, no GPU) and checked against the definition for EVERY step of random chains.

The property the kernel relies on (a true divisor can never be rejected): with
    T(k) = high 32 bits of ((N * g_k^-1 - c_lo_j) << a) mod 2^64         (c_lo_j = the cofactor floor of k's segment)
and V(k) = the value the kernel compares (tracked limb + trip offset, re-based per segment), one must have
    (V(k) - T(k)) mod 2^32  in  [0, slack + 1]        for every k,            slack = steps walked - 1,
so that T(k) <= B_j (true for every divisor in segment j, B_j = its aligned cofactor window) implies
V(k) <= B_j + slack + 1 = the kernel's threshold seg_bound[j].  GPU tests can only plant a handful of divisors; this
checks the inequality itself on hundreds of thousands of steps, for plans with ragged last trips and unequal segments.

Every chain is walked twice: by the Python transcription below and by the kernel's OWN arithmetic
(qimcifa_b200/csrc/qcf_filter_math.cuh, the header qcf_filter_kernel is compiled from, built for the host by
tests/device_math_host.cpp); the two must agree on every step, so the properties hold for the product source."""
import ctypes
import os
import random
import subprocess

from qimcifa_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_HOST = None


def host_lib():
    """tests/_build/libdevice_math_host.so: csrc/qcf_device.cuh + csrc/qcf_filter_math.cuh compiled with g++."""
    global _HOST
    if _HOST is None:
        out = os.path.join(ROOT, "tests", "_build", "libdevice_math_host.so")
        src = os.path.join(ROOT, "tests", "device_math_host.cpp")
        csrc = os.path.join(ROOT, "qimcifa_b200", "csrc")
        deps = [src, os.path.join(csrc, "qcf_device.cuh"), os.path.join(csrc, "qcf_filter_math.cuh")]
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas",
                                   "-I", csrc, src, "-o", out])
        _HOST = ctypes.CDLL(out)
        u32, u64, p32 = ctypes.c_uint32, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32)
        _HOST.h_filter_chain.restype = u32
        _HOST.h_filter_chain.argtypes = [ctypes.c_int, u64, u64, u32, u32, u32, u64, u32, u32, p32, p32, p32]
    return _HOST


def compiled_chain(N, g0, P, tp, a, D, U, seg_trips, c_lo_launch, seg_shift, slack):
    """The same walk by the kernel's own source (host build)."""
    J, steps = len(seg_trips), sum(seg_trips) * U
    out = (ctypes.c_uint32 * steps)()
    got_u = host_lib().h_filter_chain(D, N & M64, g0 & M64, P, tp, a, c_lo_launch & M64, slack, J,
                                      (ctypes.c_uint32 * J)(*seg_trips), (ctypes.c_uint32 * J)(*seg_shift), out)
    assert got_u == U
    return list(out)

M32, M64 = (1 << 32) - 1, (1 << 64) - 1


def kernel_chain(N, g0, P, tp, a, D, U, seg_trips, c_lo_launch, seg_shift, slack):
    """Transcription of the chain walk of qcf_filter_kernel: V(k) for every step walked (whole trips)."""
    x = pow(g0 & M64, -1, 1 << 64)
    nx = (N * x) & M64
    y = ((P * x) << tp) & M64
    z = ((((nx - c_lo_launch) << a) & M64) >> 32) + slack & M32
    z1 = (-(((nx << a) & M64) * y)) & M64
    d1 = d2 = d3 = 0
    if D == 1:
        d1 = z1 >> 32
    else:
        z2 = (-(z1 * y)) & M64
        if D == 2:
            d1 = ((z1 + z2) & M64) >> 32
            d2 = ((z2 + z2) & M64) >> 32
        else:
            z3 = (-(z2 * y)) & M64
            d1 = ((z1 + z2 + z3) & M64) >> 32
            d2 = ((2 * z2 + 6 * z3) & M64) >> 32
            d3 = ((6 * z3) & M64) >> 32
    s = [(u * d1 + (u * (u - 1) // 2) * d2 + (u * (u - 1) * (u - 2) // 6) * d3) & M32 for u in range(U)]
    f3 = (U * d3) & M32
    out = []
    for sg, trips in enumerate(seg_trips):
        z = (z + seg_shift[sg]) & M32
        for _ in range(trips):
            out += [(z + s[u]) & M32 for u in range(U)]
            e = (U * d2 + (U * (U - 1) // 2) * d3) & M32
            z = (z + U * d1 + (U * (U - 1) // 2) * d2 + (U * (U - 1) * (U - 2) // 6) * d3) & M32
            d1 = (d1 + e) & M32
            if D >= 3:
                d2 = (d2 + f3) & M32
            if D >= 2:
                for u in range(1, U):
                    s[u] = (s[u] + u * e + ((u * (u - 1) // 2) * f3 if D >= 3 else 0)) & M32
    return out


WHEEL11 = [v for v in range(1, 2310) if v % 2 and v % 3 and v % 5 and v % 7 and v % 11]


class Plan:
    """The planner's launch for the level-5 periods inside [v_lo, v_hi], with the windows recomputed from their definition."""

    def __init__(self, N, v_lo, v_hi, forced=True):
        pl = abi.plan_filter(N, 5, v_lo, v_hi, forced=forced)
        self.ok = pl is not None
        if not self.ok:
            return
        self.N, self.P = N, pl.period
        self.D, self.U, self.tp, self.a = pl.degree, pl.unroll, pl.tprime, pl.align
        self.K, self.trips, self.J, self.slack = pl.steps, pl.trips, pl.n_seg, pl.slack
        self.seg_trips = list(pl.seg_trips[: self.J])
        self.seg_shift = list(pl.seg_shift[: self.J])
        self.seg_bound = list(pl.seg_bound[: self.J])
        self.m_lo, self.m_end = abi.limbs_to_int(pl.m_lo_le), abi.limbs_to_int(pl.m_end_le)
        self.c_lo = pl.c_lo
        # structural invariants of a plan
        assert self.P == 2310 and self.U == (16 if self.D >= 3 else 32)
        assert sum(self.seg_trips) == self.trips and self.trips * self.U >= self.K > (self.trips - 1) * self.U
        assert self.slack == self.trips * self.U - 1
        assert self.m_end - self.m_lo == self.K << self.tp
        assert self.m_lo * self.P >= v_lo and self.m_end * self.P <= v_hi + 1
        t = self.tp + 1
        assert self.a + 2 * t >= 32 and 64 - self.a <= (self.D + 1) * t
        # the windows, from their definition
        self.seg_of, self.c_los, self.Bs = [], [], []
        p0 = self.m_lo
        for j in range(self.J):
            p1 = p0 + ((self.seg_trips[j] * self.U) << self.tp)
            chj, clj = N // (p0 * self.P if p0 else 1), N // (p1 * self.P)
            assert (chj - clj).bit_length() <= 64 - self.a, "a segment's cofactor window must fit the checked bits"
            self.c_los.append(clj)
            self.Bs.append(((chj - clj) << self.a) >> 32)
            assert self.seg_bound[j] == self.Bs[j] + self.slack + 1 < M32
            self.seg_of += [j] * (self.seg_trips[j] * self.U)
            p0 = p1
        assert self.c_lo == (N // (p0 * self.P)) & M64

    def walk(self, c, r):
        g0 = (self.m_lo + c) * self.P + r
        args = (self.N, g0, self.P, self.tp, self.a, self.D, self.U, self.seg_trips, self.c_lo, self.seg_shift, self.slack)
        v = kernel_chain(*args)
        assert compiled_chain(*args) == v, "the kernel's source and its transcription disagree"
        return g0, v


def _random_interval(rnd, N):
    top = int(N ** 0.5) >> rnd.randrange(0, 7)
    span = rnd.choice((1 << 22, 1 << 26, 1 << 30, 1 << 34, 1 << 38)) * rnd.randrange(3, 40)
    v_hi = max(top, 10 ** 6)
    v_lo = max(v_hi - span, v_hi // 2 + 1)  # one octave at most, as run_values cuts it
    return v_lo, v_hi


def test_tracked_limb_brackets_the_true_limb(monkeypatch):
    rnd = random.Random(2024)
    checked = plans = ragged = unequal = 0
    degrees = set()
    for trial in range(500):
        nbits = rnd.choice((64, 80, 96, 112, 128))
        N = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
        v_lo, v_hi = _random_interval(rnd, N)
        # the planner's own choice is degree 2 or 3 almost everywhere: force the others now and then
        abi.set_option("plan_force", "%d,0,%d,%d" % (rnd.choice((0, 0, 1, 2, 3)), rnd.choice((0, 0, 1, 4, 32)), rnd.choice((0, 1, 2))))
        pl = Plan(N, v_lo, v_hi, forced=rnd.random() < 0.5)
        if not pl.ok:
            continue
        plans += 1
        degrees.add(pl.D)
        ragged += pl.K % pl.U != 0
        unequal += len(set(pl.seg_trips)) > 1
        for _ in range(2):
            g0, V = pl.walk(rnd.randrange(0, 1 << pl.tp), rnd.choice(WHEEL11))
            ks = range(pl.K) if pl.K <= 600 else sorted(set(rnd.randrange(0, pl.K) for _ in range(600)) | {0, pl.K - 1})
            for k in ks:
                g = g0 + k * (pl.P << pl.tp)
                q = (N * pow(g & M64, -1, 1 << 64)) & M64
                T = ((((q - pl.c_los[pl.seg_of[k]]) & M64) << pl.a) & M64) >> 32
                assert 0 <= (V[k] - T) & M32 <= pl.slack + 1, (trial, pl.D, pl.tp, pl.a, pl.seg_trips, k)
                checked += 1
    assert plans > 150 and checked > 100000 and ragged > 50 and unequal > 30 and degrees == {1, 2, 3}, (plans, checked, ragged, unequal, degrees)


def test_true_divisors_pass_the_threshold(monkeypatch):
    """End to end on the model: N = g * C for a guess g of a planned launch; the kernel's compared value is within the
    threshold of g's segment, wherever g sits (first/last step, ragged last trip, every segment)."""
    rnd = random.Random(77)
    passed = 0
    seen_segments = set()
    for trial in range(1500):
        # the planner's own choice is a handful of segments: force many, and the unpadded variant, now and then
        abi.set_option("plan_force", "0,0,%d,%d" % (rnd.choice((0, 0, 8, 32)), rnd.choice((0, 0, 1, 2))))
        gbits = rnd.choice((34, 40, 48, 56, 63))
        g_mid = rnd.randrange(1 << (gbits - 1), 1 << gbits)
        span = rnd.choice((1 << 24, 1 << 28, 1 << 32, 1 << 36)) * rnd.randrange(3, 40)
        v_hi = g_mid + span // 2
        v_lo = max(g_mid - span // 2, v_hi // 2 + 1)
        # the planner sees N only through its size and the interval, so plan for a stand-in first to learn the launch
        # shape, pick a guess of that launch, then build the real N around it and re-plan
        C = rnd.randrange(v_hi, min(v_hi << rnd.randrange(1, 8), 1 << 64)) | 1
        shape = Plan(g_mid * C, v_lo, v_hi)
        if not shape.ok:
            continue
        c, r = rnd.randrange(0, 1 << shape.tp), rnd.choice(WHEEL11)
        kstar = rnd.choice((0, shape.K - 1, rnd.randrange(0, shape.K)))
        g = (shape.m_lo + c + (kstar << shape.tp)) * shape.P + r
        N = g * C
        if N.bit_length() > 128:
            continue
        pl = Plan(N, v_lo, v_hi)
        if not pl.ok or not (pl.m_lo * pl.P < g < pl.m_end * pl.P):
            continue
        m = g // pl.P - pl.m_lo
        c, k = m & ((1 << pl.tp) - 1), m >> pl.tp
        g0, V = pl.walk(c, r)
        assert g0 + k * (pl.P << pl.tp) == g
        assert V[k] <= pl.seg_bound[pl.seg_of[k]], (trial, pl.D, pl.tp, pl.a, pl.seg_trips, k, V[k])
        seen_segments.add((pl.J, pl.seg_of[k]))
        passed += 1
    assert passed > 400 and len(seen_segments) > 20, (passed, len(seen_segments))


def _q2g2(n, g0, g1, ninv, ln):
    """Transcription of qcf_divides_q2g2 (qimcifa_b200/csrc/qcf_exact_chain.cu): two-limb guess, two quotient words."""
    nl = [(n >> (32 * i)) & M32 for i in range(4)]
    n3 = nl[3] if ln > 3 else 0
    q0 = (nl[0] * ninv) & M32
    a0 = q0 * g0 + nl[0]
    b0 = (q0 * g1 + nl[1] + (a0 >> 32)) & M64
    t0 = (((n3 << 32) | nl[2]) + (b0 >> 32)) & M64  # may wrap for a non-divisor of a 4-limb N
    q1 = ((b0 & M32) * ninv) & M32
    a1 = q1 * g0 + (b0 & M32)
    b1 = (q1 * g1 + (t0 & M32) + (a1 >> 32)) & M64
    t1 = (t0 >> 32) + (b1 >> 32)
    return (b1 & M32) == g0 and t1 == g1


def test_two_limb_reduction_model():
    rnd = random.Random(31337)
    for trial in range(20000):
        ln = rnd.choice((3, 4))
        gbits = rnd.randrange(33, 65)
        g = rnd.randrange(1 << (gbits - 1), 1 << gbits) | 1
        if rnd.random() < 0.5:
            c = rnd.randrange(1, 1 << min(64, 32 * ln - gbits))  # cofactor below 2^64: the QN = 2 precondition
            n = g * c
        else:
            n = rnd.randrange(1 << (32 * ln - 8), 1 << (32 * ln))
            if n // g >= 1 << 64:
                n = g * ((1 << 64) - 1) + rnd.randrange(0, g)
        if n.bit_length() > 32 * ln or n // g >= 1 << 64 or n == 0:
            continue
        ninv = (-pow(g & M32, -1, 1 << 32)) & M32
        assert _q2g2(n, g & M32, g >> 32, ninv, ln) == (n % g == 0), (n, g, ln)
    # extreme operands: N and g at the top of their ranges
    for g in ((1 << 64) - 1, (1 << 64) - 59, (1 << 33) + 1):
        for c in (1, 3, (1 << 64) - 1, (1 << 63) + 1):
            n = g * c
            if n.bit_length() <= 128:
                ninv = (-pow(g & M32, -1, 1 << 32)) & M32
                assert _q2g2(n, g & M32, g >> 32, ninv, 4)
                assert not _q2g2(n + 2 if (n + 2) % g else n + 4, g & M32, g >> 32, ninv, 4) or (n + 2).bit_length() > 128
