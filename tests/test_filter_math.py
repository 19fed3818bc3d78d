"""CPU-only: the arithmetic of the filter kernel (qimcifa_b200/csrc/qcf_filter.cu + plan_filter in qcf_api.cu), restated
in exact Python integers and checked against the definition for EVERY step of random chains.

The property the kernel relies on (a true divisor can never be rejected): with
    T(k) = high 32 bits of ((N * g_k^-1 - c_lo_j) << a) mod 2^64         (c_lo_j = the cofactor floor of k's segment)
and V(k) = the value the kernel compares (tracked limb + trip offset, re-based per segment), one must have
    (V(k) - T(k)) mod 2^32  in  [0, K]        for every k,
so that T(k) <= B_j implies V(k) <= B_j + K = the kernel's threshold.  GPU tests can only plant a handful of divisors;
this checks the inequality itself on millions of steps."""
import random

M32, M64 = (1 << 32) - 1, (1 << 64) - 1


def unroll_for(D):
    return 16 if D >= 3 else 32


def kernel_chain(N, g0, P, tp, a, D, K, n_seg, c_lo_launch, seg_shift, slack):
    """Transcription of the chain walk of qcf_filter_kernel: yields V(k) for k = 0..K-1."""
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
    U = unroll_for(D)
    s = [(u * d1 + (u * (u - 1) // 2) * d2 + (u * (u - 1) * (u - 2) // 6) * d3) & M32 for u in range(U)]
    f3 = (U * d3) & M32
    out = []
    seg_steps = K // n_seg
    for sg in range(n_seg):
        z = (z + seg_shift[sg]) & M32
        for _ in range(seg_steps // U):
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


def plan_segments(N, P, m_lo, tp, K, n_seg, a):
    """Transcription of the per-segment windows of plan_filter: -> (c_lo_launch, seg_shift, seg_B, seg_c_lo)."""
    seg_periods = (K // n_seg) << tp
    m_end = m_lo + seg_periods * n_seg
    c_lo_launch = N // (m_end * P)
    shifts, Bs, clos = [], [], []
    prev = 0
    for j in range(n_seg):
        p0 = m_lo + seg_periods * j
        p1 = p0 + seg_periods
        chj = N // (p0 * P if p0 * P else 1)
        clj = N // (p1 * P)
        Bs.append(((chj - clj) << a) >> 32)
        delta = (((clj - c_lo_launch) & M64) << a) & M64
        dhi = delta >> 32
        shifts.append((-dhi) & M32 if j == 0 else (prev - dhi) & M32)
        prev = dhi
        clos.append(clj)
    return c_lo_launch, shifts, Bs, clos


def test_tracked_limb_brackets_the_true_limb():
    rnd = random.Random(2024)
    P = 2310
    checked = 0
    for trial in range(1500):
        D = rnd.choice((1, 2, 3))
        tp = rnd.randrange(3, 30)
        t = tp + 1
        w = min(64, (D + 1) * t)
        a = 64 - w
        if a + 2 * t < 32:
            continue
        U = unroll_for(D)
        n_seg = rnd.choice((1, 4, 16))
        K = U * n_seg * rnd.randrange(1, 4)
        nbits = rnd.choice((64, 96, 112, 128))
        N = rnd.randrange(1 << (nbits - 1), 1 << nbits) | 1
        # a launch somewhere at or below sqrt(N)
        top = int(N ** 0.5) >> rnd.randrange(0, 6)
        m_lo = max(1, top // P - (K << tp) - rnd.randrange(0, 1000))
        c_lo_launch, shifts, Bs, clos = plan_segments(N, P, m_lo, tp, K, n_seg, a)
        dC0 = N // (m_lo * P) - clos[0]
        if dC0.bit_length() > w or ((dC0 << a) >> 32) + K >= M32:
            continue  # the planner rejects such a plan
        r = rnd.choice([v for v in range(1, P) if v % 2 and v % 3 and v % 5 and v % 7 and v % 11])
        c = rnd.randrange(0, 1 << tp)
        g0 = (m_lo + c) * P + r
        V = kernel_chain(N, g0, P, tp, a, D, K, n_seg, c_lo_launch & M64, shifts, K - 1)
        seg_steps = K // n_seg
        for k in range(K):
            g = g0 + k * (P << tp)
            q = (N * pow(g & M64, -1, 1 << 64)) & M64
            j = k // seg_steps
            T = ((((q - clos[j]) & M64) << a) & M64) >> 32
            diff = (V[k] - T) & M32
            assert 0 <= diff <= K, (trial, D, tp, a, n_seg, K, k, diff)
            checked += 1
            # and the window really contains every cofactor of this segment: if g | N' for N' = g * C with C in
            # [c_lo_j, c_hi_j] then T <= B_j -- spot-check with the segment's own cofactor bounds
        # exact divisors: plant C so that N2 = g * C has the same plan-relevant quantities is not possible in general;
        # instead check monotonic coverage of the segment windows
        for j in range(n_seg - 1):
            assert clos[j] >= clos[j + 1]
    assert checked > 100000


def test_true_divisors_pass_the_threshold():
    """End to end on the model: N = g * C for a guess g of the chain; the kernel's compared value is within its threshold."""
    rnd = random.Random(77)
    P = 2310
    wheel = [v for v in range(1, P) if v % 2 and v % 3 and v % 5 and v % 7 and v % 11]
    passed = 0
    for trial in range(2000):
        D = rnd.choice((1, 2, 3))
        tp = rnd.randrange(6, 26)
        t = tp + 1
        w = min(64, (D + 1) * t)
        a = 64 - w
        if a + 2 * t < 32:
            continue
        U = unroll_for(D)
        n_seg = rnd.choice((1, 4, 16))
        K = U * n_seg * rnd.randrange(1, 3)
        gbits = rnd.choice((40, 48, 56, 63))
        m_lo = rnd.randrange(1 << (gbits - 13), 1 << (gbits - 12))
        c = rnd.randrange(0, 1 << tp)
        r = rnd.choice(wheel)
        kstar = rnd.randrange(0, K)
        g0 = (m_lo + c) * P + r
        g = g0 + kstar * (P << tp)
        C = rnd.randrange(g, min(g << rnd.randrange(1, 10), 1 << 64)) | 1  # cofactor above the guess: g is the smaller divisor
        N = g * C
        if N.bit_length() > 128:
            continue
        c_lo_launch, shifts, Bs, clos = plan_segments(N, P, m_lo, tp, K, n_seg, a)
        dC0 = N // (m_lo * P) - clos[0]
        if dC0.bit_length() > w or ((dC0 << a) >> 32) + K >= M32:
            continue
        V = kernel_chain(N, g0, P, tp, a, D, K, n_seg, c_lo_launch & M64, shifts, K - 1)
        j = kstar // (K // n_seg)
        assert V[kstar] <= Bs[j] + K, (trial, D, tp, a, n_seg, K, kstar, V[kstar], Bs[j])
        passed += 1
    assert passed > 500


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
