"""CPU-only: the arithmetic of the filter kernel (qimcifa_b200/csrc/qcf_filter.cu), restated in exact Python integers,
run on the launch parameters the REAL planner produces (plan_filter in qcf_api.cu, through the qcf_plan_filter hook --
host code, no GPU) and checked against the definition for EVERY step of random chains.

The property the kernel relies on (a true divisor can never be rejected): with
    T(k) = high 32 bits of ((N * g_k^-1 - l_j(k)) << a) mod 2^64
(l_j(k) = beta_j - lambda_j * (k - ka_j): the reference LINE of k's segment, which follows the cofactor down the chain;
lambda_j = 0, a constant floor, when the drift is switched off) and V(k) = the value the kernel compares (tracked limb +
trip offset, re-based per segment), one must have
    (V(k) - T(k)) mod 2^32  in  [0, slack]     for every k,     slack = (steps walked - 1) + (segments - 1),
so that T(k) <= B_j (true for every divisor in segment j, B_j = its aligned window) implies V(k) <= B_j + slack = the
kernel's threshold seg_bound[j].  The windows are checked from their definition too: for every step k of a segment the line
lies below f(k+1) = N / u_{k+1} and at most W_j below f(k) = N / u_k (u_k = base + k * stride), and every guess of step k
lies strictly between u_k and u_{k+1}: so the cofactor of a divisor at step k is within [l_j(k), l_j(k) + W_j].  GPU tests can only plant a handful of divisors; this
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
        deps = [src, os.path.join(csrc, "qcf_device.cuh"), os.path.join(csrc, "qcf_filter_math.cuh"), os.path.join(csrc, "qcf_trip.h")]
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas",
                                   "-I", csrc, src, "-o", out])
        _HOST = ctypes.CDLL(out)
        u32, u64, p32 = ctypes.c_uint32, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32)
        _HOST.h_filter_chain.restype = u32
        _HOST.h_filter_chain.argtypes = [ctypes.c_int, u64, u64, u32, u32, u32, u64, u64, u32, u32, p32, p32, p32, p32, u32, u64, u32, u32, u64, p32, p32]
        _HOST.h_filter_chain_pk2.restype = u32
        _HOST.h_filter_chain_pk2.argtypes = [u64, u64, u32, u32, u32, u64, u64, u32, u32, p32, p32, p32, p32, p32, u32, u64, u32, u32, u64, p32, u32, p32, p32]
        _HOST.h_filter_chain16.restype = u32
        _HOST.h_filter_chain16.argtypes = [u64, u64, u32, u32, u32, u64, u64, u32, u32, p32, p32, p32, p32, p32, u32, u64, u32, u32, u64, p32, p32, p32]
        _HOST.h_filter_flagged_min.restype = u32
        _HOST.h_filter_flagged_min.argtypes = [u32, u32, u32, u32, u32]
    return _HOST


def compiled_chain(N, g0, P, tp, a, D, U, seg_trips, c_lo_launch, mu0, seg_shift, seg_nu, slack, rho):
    """The same walk by the kernel's own source (host build).  rho = None or (c, r, pinv32, sigma0, seg_kappa)."""
    J, steps = len(seg_trips), sum(seg_trips) * U
    out = (ctypes.c_uint32 * steps)()
    c, r, pinv32, sigma0, kappa = rho if rho else (0, 0, 0, 0, [0] * J)
    got_u = host_lib().h_filter_chain(D, N & M64, g0 & M64, P, tp, a, c_lo_launch & M64, mu0 & M64, slack, J,
                                      (ctypes.c_uint32 * J)(*seg_trips), (ctypes.c_uint32 * J)(*seg_shift),
                                      (ctypes.c_uint32 * J)(*[v & M32 for v in seg_nu]), (ctypes.c_uint32 * J)(*[v >> 32 for v in seg_nu]),
                                      1 if rho else 0, c, r, pinv32, sigma0, (ctypes.c_uint32 * J)(*kappa), out)
    assert got_u == U
    return list(out)

def compiled_chain16(N, g0, P, tp, a, U, seg_trips, c_lo_launch, mu0, seg_shift, seg_bound, seg_nu, slack, rho):
    """The degree-1 walk as the kernel makes it (packed 16-bit trips): (full-width value per step, replay flag per trip)."""
    J, trips = len(seg_trips), sum(seg_trips)
    out, flag = (ctypes.c_uint32 * (trips * U))(), (ctypes.c_uint32 * trips)()
    c, r, pinv32, sigma0, kappa = rho if rho else (0, 0, 0, 0, [0] * J)
    arr = lambda v: (ctypes.c_uint32 * J)(*v)  # noqa: E731
    got_u = host_lib().h_filter_chain16(N & M64, g0 & M64, P, tp, a, c_lo_launch & M64, mu0 & M64, slack, J, arr(seg_trips), arr(seg_shift),
                                        arr(seg_bound), arr([v & M32 for v in seg_nu]), arr([v >> 32 for v in seg_nu]),
                                        1 if rho else 0, c, r, pinv32, sigma0, arr(kappa), out, flag)
    assert got_u == U
    return list(out), list(flag)


def compiled_chain_pk2(N, g0, P, tp, a, U, seg_trips, c_lo_launch, mu0, seg_shift, seg_bound, seg_nu, slack, rho, run_trips):
    """The degree-2 walk in the packed form (sub-trips of 16 in groups of 8): (full-width value per step, replay flag per trip)."""
    J, trips = len(seg_trips), sum(seg_trips)
    out, flag = (ctypes.c_uint32 * (trips * U))(), (ctypes.c_uint32 * trips)()
    c, r, pinv32, sigma0, kappa = rho if rho else (0, 0, 0, 0, [0] * J)
    arr = lambda v: (ctypes.c_uint32 * J)(*v)  # noqa: E731
    got_u = host_lib().h_filter_chain_pk2(N & M64, g0 & M64, P, tp, a, c_lo_launch & M64, mu0 & M64, slack, J, arr(seg_trips), arr(seg_shift),
                                          arr(seg_bound), arr([v & M32 for v in seg_nu]), arr([v >> 32 for v in seg_nu]),
                                          1 if rho else 0, c, r, pinv32, sigma0, arr(kappa), run_trips, out, flag)
    assert got_u == U
    return list(out), list(flag)


M32, M64 = (1 << 32) - 1, (1 << 64) - 1


def rho32_of(c, r, pinv32, tp):
    """Transcription of qcf_filter_rho32: the chain's offset, normalised to 32 bits (slightly under)."""
    return min(M32, ((c << 32) | ((r * pinv32) & M32)) >> tp)


def kernel_chain(N, g0, P, tp, a, D, U, seg_trips, c_lo_launch, mu0, seg_shift, seg_nu, slack, rho):
    """Transcription of the chain walk of qcf_filter_kernel: V(k) for every step walked (whole trips)."""
    x = pow(g0 & M64, -1, 1 << 64)
    nx = (N * x) & M64
    y = ((P * x) << tp) & M64
    rho32, corr, kappa = 0, 0, [0] * len(seg_trips)
    if rho:
        c, r, pinv32, sigma0, kappa = rho
        rho32 = rho32_of(c, r, pinv32, tp)
        corr = rho32 * (sigma0 >> 32) + ((rho32 * (sigma0 & M32)) >> 32) + 1  # qcf_filter_rho_corr
    z = ((((nx - c_lo_launch + corr) << a) & M64) >> 32) + slack & M32
    z1 = (-(((nx << a) & M64) * y)) & M64
    d2 = d3 = 0
    first = (z1 + mu0) & M64  # the first difference, both limbs: the drift slope of segment 0 is part of it
    if D >= 2:
        z2 = (-(z1 * y)) & M64
        first = (first + z2) & M64
        if D == 2:
            d2 = ((z2 + z2) & M64) >> 32
        else:
            z3 = (-(z2 * y)) & M64
            first = (first + z3) & M64
            d2 = ((2 * z2 + 6 * z3) & M64) >> 32
            d3 = ((6 * z3) & M64) >> 32
    d1, elo = first >> 32, first & M32
    s = [(u * d1 + (u * (u - 1) // 2) * d2 + (u * (u - 1) * (u - 2) // 6) * d3) & M32 for u in range(U)]
    f3 = (U * d3) & M32
    out = []
    for sg, trips in enumerate(seg_trips):
        z = (z + seg_shift[sg]) & M32
        if sg > 0:  # the slope of the reference line changes: the 64-bit first difference moves by nu, carry included
            z = (z - ((rho32 * kappa[sg]) >> 32)) & M32  # ... and the chain's offset correction shrinks
            lo = elo + (seg_nu[sg] & M32)
            dd = ((seg_nu[sg] >> 32) + (lo >> 32)) & M32
            elo = lo & M32
            d1 = (d1 + dd) & M32
            s = [(s[u] + u * dd) & M32 for u in range(U)]
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
        self.packed = bool(pl.packed)
        self.K, self.trips, self.J, self.slack = pl.steps, pl.trips, pl.n_seg, pl.slack
        self.seg_trips = list(pl.seg_trips[: self.J])
        self.seg_shift = list(pl.seg_shift[: self.J])
        self.seg_bound = list(pl.seg_bound[: self.J])
        self.drift, self.mu0 = bool(pl.drift), pl.mu0
        self.seg_nu = [(pl.seg_nu_hi[j] << 32) | pl.seg_nu_lo[j] for j in range(self.J)]
        self.lam = list(pl.seg_lambda[: self.J])
        self.ka = list(pl.seg_ka[: self.J])
        self.beta = [abi.limbs_to_int(pl.seg_beta_le[j]) for j in range(self.J)]
        self.W = list(pl.seg_window[: self.J])
        self.rho, self.pinv32, self.sigma0 = bool(pl.rho), pl.pinv32, pl.sigma0
        self.sigma, self.kappa = list(pl.seg_sigma[: self.J]), list(pl.seg_kappa[: self.J])
        self.over = 2 * (self.J - 1) if self.rho else 0  # the per-segment correction may over-subtract by 2: in the thresholds
        self.m_lo, self.m_end = abi.limbs_to_int(pl.m_lo_le), abi.limbs_to_int(pl.m_end_le)
        self.c_lo = pl.c_lo
        # structural invariants of a plan
        assert self.P == 2310 and self.U == {1: 32, 2: 32, 3: 16}[self.D]
        assert sum(self.seg_trips) == self.trips and self.trips * self.U >= self.K > (self.trips - 1) * self.U
        rho_loss = ((1 << (self.a - 32)) + 1 if self.a > 32 else 2) if self.rho else 0
        assert self.slack == self.trips * self.U - 1 + (self.J - 1) * (1 + rho_loss)
        assert self.m_end - self.m_lo == self.K << self.tp
        assert self.m_lo * self.P >= v_lo and self.m_end * self.P <= v_hi + 1
        t = self.tp + 1
        assert self.a + 2 * t >= 32 and 64 - self.a <= (self.D + 1) * t
        assert self.mu0 == (self.lam[0] << self.a) & M64 and self.c_lo == self.line(0, 0) & M64
        assert self.drift or not any(self.lam)
        # the lines and windows, from their definition: u_k = base + k * stride, f(k) = N / u_k; every guess of step k lies
        # strictly between u_k and u_(k+1), so a divisor's cofactor lies strictly between f(k+1) and f(k)
        base, S = self.m_lo * self.P, self.P << self.tp
        self.seg_of, self.seg_k0 = [], []
        k0 = 0
        for j in range(self.J):
            k1 = k0 + self.seg_trips[j] * self.U
            self.seg_k0.append(k0)
            assert self.W[j].bit_length() <= 64 - self.a, "a segment's window must fit the checked bits"
            assert self.seg_bound[j] == ((self.W[j] << self.a) >> 32) + self.slack + self.over < M32
            ks = range(k0, k1) if k1 - k0 <= 400 else sorted({k0, k0 + 1, k1 - 2, k1 - 1, self.ka[j], self.ka[j] + 1} | {random.randrange(k0, k1) for _ in range(60)})
            for k in ks:
                if self.rho:
                    # every chain has its own reference: check it on the real-valued cofactor N/g of sampled chains
                    for c, r in ((0, 1), ((1 << self.tp) - 1, self.P - 1), (random.randrange(1 << self.tp), random.choice(WHEEL11))):
                        g = base + k * S + c * self.P + r
                        ref = self.ref(j, k, c, r)
                        assert ref <= N // g, ("the reference must lie below the cofactor", j, k, c, r)
                        assert -(-N // g) - ref <= self.W[j], ("window too small", j, k, c, r)
                    continue
                line = self.line(j, k)
                assert line <= N // (base + (k + 1) * S), ("the line must lie below the cofactors of step k", j, k)  # floor(f(k+1)) < C
                assert -(-N // (base + k * S)) - 1 - line <= self.W[j], ("window too small", j, k)                    # C < f(k): C <= ceil(f(k)) - 1
            if j:
                # re-basing constants between consecutive lines at the segment's first step
                delta = ((self.line(j - 1, k0) - self.line(j, k0)) << self.a) & M64
                assert self.seg_shift[j] == delta >> 32 and self.seg_nu[j] == ((self.lam[j] - self.lam[j - 1]) << self.a) & M64
                if self.rho:
                    assert self.sigma[j] <= self.sigma[j - 1] and self.kappa[j] == ((self.sigma[j - 1] - self.sigma[j]) << self.a) >> 32
            else:
                assert self.seg_shift[0] == 0 and self.seg_nu[0] == 0 and (not self.rho or self.sigma[0] == self.sigma0)
            self.seg_of += [j] * (k1 - k0)
            k0 = k1

    def line(self, j, k):
        """l_j(k): the reference line of segment j at chain step k."""
        return self.beta[j] - self.lam[j] * (k - self.ka[j])

    def ref(self, j, k, c, r):
        """What chain (c, r) compares the quotient against at step k of segment j: the line, lowered by the chain's own
        offset correction A_j (exact integers, as the kernel's arithmetic produces them up to the tracked-limb roundings)."""
        if not self.rho:
            return self.line(j, k)
        rho32 = rho32_of(c, r, self.pinv32, self.tp)
        A = ((rho32 * self.sigma[0]) >> 32) + 1
        for i in range(1, j + 1):
            A -= (rho32 * (self.sigma[i - 1] - self.sigma[i])) >> 32
        return self.line(j, k) - A

    def walk(self, c, r):
        g0 = (self.m_lo + c) * self.P + r
        rho = (c, r, self.pinv32, self.sigma0, self.kappa) if self.rho else None
        args = (self.N, g0, self.P, self.tp, self.a, self.D, self.U, self.seg_trips, self.c_lo, self.mu0, self.seg_shift, self.seg_nu, self.slack, rho)
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
    checked = plans = ragged = unequal = drifting = corrected = 0
    degrees = set()
    for trial in range(500):
        nbits = rnd.choice((64, 80, 96, 112, 128))
        N = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
        v_lo, v_hi = _random_interval(rnd, N)
        # the planner's own choice is degree 2 or 3 almost everywhere: force the others now and then
        abi.set_option("plan_force", "%d,0,%d,%d" % (rnd.choice((0, 0, 1, 2, 3)), rnd.choice((0, 0, 1, 4, 32)), rnd.choice((0, 1, 2))))
        abi.set_option("plan_drift", rnd.choice((0, 1, 2, 2)))
        pl = Plan(N, v_lo, v_hi, forced=rnd.random() < 0.5)
        if not pl.ok:
            continue
        plans += 1
        drifting += pl.drift and any(pl.lam)
        corrected += pl.rho
        degrees.add(pl.D)
        ragged += pl.K % pl.U != 0
        unequal += len(set(pl.seg_trips)) > 1
        for _ in range(2):
            c, r = rnd.randrange(0, 1 << pl.tp), rnd.choice(WHEEL11)
            g0, V = pl.walk(c, r)
            ks = range(pl.K) if pl.K <= 600 else sorted(set(rnd.randrange(0, pl.K) for _ in range(600)) | {0, pl.K - 1})
            for k in ks:
                g = g0 + k * (pl.P << pl.tp)
                q = (N * pow(g & M64, -1, 1 << 64)) & M64
                T = ((((q - pl.ref(pl.seg_of[k], k, c, r)) & M64) << pl.a) & M64) >> 32
                assert 0 <= (V[k] - T) & M32 <= pl.slack + pl.over, (trial, pl.D, pl.tp, pl.a, pl.seg_trips, k, pl.rho)
                checked += 1
    assert plans > 150 and checked > 100000 and ragged > 50 and unequal > 30 and drifting > 80 and corrected > 50 and degrees == {1, 2, 3}, (
        plans, checked, ragged, unequal, drifting, corrected, degrees)


def test_true_divisors_pass_the_threshold(monkeypatch):
    """End to end on the model: N = g * C for a guess g of a planned launch; the kernel's compared value is within the
    threshold of g's segment, wherever g sits (first/last step, ragged last trip, every segment)."""
    rnd = random.Random(77)
    passed = 0
    seen_segments = set()
    for trial in range(1500):
        # the planner's own choice is a handful of segments: force many, and the unpadded variant, now and then
        abi.set_option("plan_force", "0,0,%d,%d" % (rnd.choice((0, 0, 8, 32)), rnd.choice((0, 0, 1, 2))))
        abi.set_option("plan_drift", rnd.choice((0, 1, 2, 2, 2)))
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


def test_packed_trip_flags_every_survivor(monkeypatch):
    """Degree 1 walks packed 16-bit trips (qcf_filter_trip16): two guesses per add+min on the high halves of the tracked limb;
    a flagged trip is replayed at full width.  On real plans: the full-width values of the packed walk are those of the plain
    walk, every trip that holds a step at or below its segment's threshold is flagged (the packed test is a NECESSARY condition
    of the full-width one, so no divisor is lost), and the flag rate stays near U * T / 2^16 (it is a filter, not a constant)."""
    rnd = random.Random(1616)
    plans = trips_seen = flagged = holders = 0
    expected = 0.0
    for trial in range(400):
        nbits = rnd.choice((96, 112, 128))
        N = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
        v_lo, v_hi = _random_interval(rnd, N)
        abi.set_option("plan_force", "1,0,%d,%d" % (rnd.choice((0, 0, 1, 4, 32)), rnd.choice((0, 1, 2))))
        abi.set_option("plan_drift", rnd.choice((0, 1, 2, 2)))
        forced = rnd.random() < 0.5
        pl = Plan(N, v_lo, v_hi, forced=forced)
        if not pl.ok:
            continue
        assert pl.D == 1 and max(pl.seg_bound) < 1 << 26
        plans += 1
        for _ in range(3):
            c, r = rnd.randrange(0, 1 << pl.tp), rnd.choice(WHEEL11)
            g0, V = pl.walk(c, r)
            rho = (c, r, pl.pinv32, pl.sigma0, pl.kappa) if pl.rho else None
            V16, flags = compiled_chain16(N, g0, pl.P, pl.tp, pl.a, pl.U, pl.seg_trips, pl.c_lo, pl.mu0, pl.seg_shift, pl.seg_bound, pl.seg_nu, pl.slack, rho)
            assert V16 == V
            for t, f in enumerate(flags):
                b = pl.seg_bound[pl.seg_of[t * pl.U]]
                holds = any(v <= b for v in V[t * pl.U:(t + 1) * pl.U])
                assert f or not holds, (trial, t, b)
                # the flag names the residue u mod 4 of every step at or below the bound: the replay looks at those steps only
                assert all((f >> (u & 3)) & 1 for u, v in enumerate(V[t * pl.U:(t + 1) * pl.U]) if v <= b), (trial, t, b, f)
                trips_seen += 1
                flagged += 1 if f else 0
                holders += holds
                expected += pl.U * ((b >> 16) + 1.5) / 65536.0
    # also a chain built to survive: a true divisor's step is flagged (end to end through the packed test)
    assert plans > 100 and trips_seen > 8000 and holders > 0, (plans, trips_seen, holders)
    assert flagged < 2.0 * expected + 50 and flagged > 0.3 * expected - 50, (flagged, expected)


def test_packed_degree2_flags_every_survivor(monkeypatch):
    """Degree 2 in the packed form: sub-trips of 16 guesses in groups of 8 whose offsets come from the group's exact base by a
    32-bit multiply-add on two packed halves (the high half picks up carries, kept within 0..4 by rounding the multiplier: its threshold is widened by 4).  On real plans
    that the planner marks `packed`: the full-width values equal the plain walk's, every trip holding a step at or below its
    segment's threshold is flagged, whatever the run length that cuts the groups (epoch budgets of 1, 3, 4, 5, 1000 trips), and
    the flag rate stays that of a filter."""
    rnd = random.Random(2222)
    plans = trips_seen = flagged = holders = 0
    expected = 0.0
    for trial in range(400):
        nbits = rnd.choice((64, 80, 96, 112))
        N = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
        v_lo, v_hi = _random_interval(rnd, N)
        abi.set_option("plan_force", "2,0,%d,%d" % (rnd.choice((0, 0, 1, 4, 32)), rnd.choice((0, 1, 2))))
        abi.set_option("plan_drift", rnd.choice((0, 1, 2, 2)))
        pl = Plan(N, v_lo, v_hi, forced=rnd.random() < 0.5)
        if not pl.ok or not pl.packed:
            continue
        assert pl.D == 2 and pl.tp + 1 <= 21 and max(pl.seg_bound) < 1 << 26
        plans += 1
        for _ in range(2):
            c, r = rnd.randrange(0, 1 << pl.tp), rnd.choice(WHEEL11)
            g0, V = pl.walk(c, r)
            rho = (c, r, pl.pinv32, pl.sigma0, pl.kappa) if pl.rho else None
            run_trips = rnd.choice((1, 3, 4, 5, 1000))
            Vp, flags = compiled_chain_pk2(N, g0, pl.P, pl.tp, pl.a, pl.U, pl.seg_trips, pl.c_lo, pl.mu0, pl.seg_shift, pl.seg_bound, pl.seg_nu, pl.slack, rho, run_trips)
            assert Vp == V, (trial, run_trips)
            for t, f in enumerate(flags):
                b = pl.seg_bound[pl.seg_of[t * pl.U]]
                holds = any(v <= b for v in V[t * pl.U:(t + 1) * pl.U])
                assert f or not holds, (trial, t, b, run_trips)
                # the flag names the residue u mod 4 of every step at or below the bound: the replay looks at those steps only
                assert all((f >> (u & 3)) & 1 for u, v in enumerate(V[t * pl.U:(t + 1) * pl.U]) if v <= b), (trial, t, b, f)
                trips_seen += 1
                flagged += 1 if f else 0
                holders += holds
                expected += pl.U * ((b >> 16) + 4.0) / 65536.0
    assert plans > 60 and trips_seen > 10000 and holders > 0, (plans, trips_seen, holders)
    assert flagged < 2.0 * expected + 50 and flagged > 0.3 * expected - 50, (flagged, expected)


def test_flagged_min_is_the_minimum_over_the_flagged_classes():
    """qcf_filter_flagged_min (the replay of a flagged packed trip): for every class mask, the minimum over the steps
    u = r (mod 4), r in the mask, of v(u) = z + u*d1 + C(u,2)*d2 modulo 2^32 -- against the step-by-step recurrence the
    kernel's queueing loop walks (v += dv, dv += d2)."""
    rnd = random.Random(4242)
    lib = host_lib()
    for trial in range(20000):
        U = rnd.choice((16, 32, 64, 128))  # one, two or four trips under one flag test
        z, d1 = rnd.randrange(1 << 32), rnd.randrange(1 << 32)
        d2 = rnd.choice((0, rnd.randrange(1 << 32), rnd.randrange(1 << 20) << 12))
        vals, v, dv = [], z, d1
        for u in range(U):
            vals.append(v)
            v = (v + dv) & M32
            dv = (dv + d2) & M32
        cls = rnd.randrange(1, 16)
        want = min(x for u, x in enumerate(vals) if (cls >> (u & 3)) & 1)
        assert lib.h_filter_flagged_min(U, z, d1, d2, cls) == want, (trial, z, d1, d2, cls)
    assert lib.h_filter_flagged_min(32, 5, 7, 9, 0) == M32  # a lane without flags never enters the queueing loop


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
