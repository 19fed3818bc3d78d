// Chain arithmetic of the filter kernel (qcf_filter.cu), kept apart from the kernel so that a host compiler can build
// it too: tests/device_math_host.cpp walks whole chains with THIS code on the plans of the real planner and
// tests/test_filter_math.py checks every step against the definition in Python integers.  See qcf_filter.cu for the
// mathematics; everything here is exact arithmetic modulo 2^64 / 2^32.
#pragma once

#include "qcf_device.cuh"

#include "qcf_trip.h" // QCF_FILTER_UNROLL_FOR: guesses per trip by degree

// Chain setup for the chain starting at guess g0 (only g0 mod 2^64 matters): x = g0^-1 mod 2^64, then the tracked high
// limb z of  R(0) = ((N*g0^-1 - ref) << al) mod 2^64  (started `slack` too high, see qcf_filter.cu) and the forward
// differences of R: the first one as a full 64-bit value d1:elo (its low limb is constant inside a segment; the drift
// slope of the segment's reference line, mu = lambda << al, is part of it), the higher ones by their high limbs (their
// low limbs are zero).
//   n64 = N mod 2^64, P = period, tp = log2 of the stride in periods, al = alignment shift, ref = the reference line's
//   value at step 0 (mod 2^64), mu0 = the slope term of segment 0.
//   corr = the chain's own offset correction (qcf_filter_rho_corr), 0 when the launch does not use it.
template <int D> __device__ __forceinline__ void qcf_filter_chain_init(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 ref, u64 mu0, u64 corr, u32 slack, u32& z, u32& d1, u32& elo, u32& d2, u32& d3)
{
    u64 x = qcf_inv32((u32)g0);
    x *= 2ull - g0 * x;
    const u64 nx = n64 * x;
    const u64 y = ((u64)P * x) << tp; // S * x
    z = (u32)(((nx - ref + corr) << al) >> 32) + slack;
    d2 = 0, d3 = 0;
    const u64 z1 = 0ull - (nx << al) * y;
    u64 first = z1 + mu0;
    if (D >= 2) {
        const u64 z2 = 0ull - z1 * y;
        first += z2;
        if (D == 2) {
            d2 = (u32)((z2 + z2) >> 32);
        } else {
            const u64 z3 = 0ull - z2 * y;
            first += z3;
            d2 = (u32)((2ull * z2 + 6ull * z3) >> 32);
            d3 = (u32)((6ull * z3) >> 32);
        }
    }
    d1 = (u32)(first >> 32);
    elo = (u32)first;
}

// The chain's offset inside its step: every chain of a launch starts in (base, base + S), chain (c, r) at rho = c*P + r.  A
// guess at step k is u_k + rho, and its cofactor lies rho * N/u^2 BELOW f(k) = N/u_k -- up to a whole step's drift, which
// is what the windows had to absorb.  Instead each chain lowers its own reference by that amount:
//   rho32 = floor(rho * 2^32 / S), slightly under (pinv32 = floor(2^32 / P));   correction = floor(rho32 * sigma / 2^32) + 1
// with sigma = the launch-uniform bound on one step's drift in the current segment (sigma0 at setup; entering segment j the
// tracked limb drops by umulhi(rho32, kappa_j), kappa_j = ((sigma_(j-1) - sigma_j) << al) >> 32).  What remains in the window
// is rho * (the CHANGE of N/u^2 over one segment).  The planner accounts for every rounding here (plan_filter, qcf_api.cu)
// and tests/test_filter_math.py checks the resulting windows on sampled chains and steps.
__device__ __forceinline__ u32 qcf_filter_rho32(u64 c, u32 r, u32 pinv32, u32 tp)
{
    const u64 v = ((c << 32) | (u64)(r * pinv32)) >> tp; // c < 2^tp, r < P: r * pinv32 < 2^32
    return (v > 0xFFFFFFFFull) ? 0xFFFFFFFFu : (u32)v;
}
__device__ __forceinline__ u64 qcf_filter_rho_corr(u32 rho32, u64 sigma)
{
    return (u64)rho32 * (u32)(sigma >> 32) + (((u64)rho32 * (u32)sigma) >> 32) + 1ull; // floor(rho32 * sigma / 2^32) + 1, sigma < 2^63
}

// Entering a segment: the reference line changes.  Its value at the segment's first step moves the tracked limb by a
// launch-uniform constant (`shift`; the carry out of the untracked low limb is covered by the slack), its slope moves the
// 64-bit first difference d1:elo by nu = (lambda_j - lambda_{j-1}) << al exactly (carry from the low limb included), and the
// trip offsets s[u] = u*d1 + ... follow d1.
template <u32 U> __device__ __forceinline__ void qcf_filter_enter_segment(u32& z, u32& d1, u32& elo, u32 (&s)[U], u32 shift, u32 nu_lo, u32 nu_hi, u32 rho32, u32 kappa)
{
    z += shift - (u32)(((u64)rho32 * kappa) >> 32);
    const u32 lo = elo + nu_lo;
    const u32 dd = nu_hi + ((lo < elo) ? 1u : 0u);
    elo = lo;
    d1 += dd;
#pragma unroll
    for (u32 u = 1; u < U; ++u) {
        s[u] += u * dd;
    }
}

// Trip-relative offsets: inside a trip starting at step k0 the tracked limb is z(k0 + u) = z + s[u],
//   s[u] = u*d1 + C(u,2)*d2 + C(u,3)*d3   (d1, d2 = the differences at k0).
template <u32 U> __device__ __forceinline__ void qcf_filter_offsets(u32 (&s)[U], u32 d1, u32 d2, u32 d3)
{
#pragma unroll
    for (u32 u = 0; u < U; ++u) {
        s[u] = u * d1 + (u * (u - 1u) / 2u) * d2 + (u * (u - 1u) * (u - 2u) / 6u) * d3;
    }
}

// Moves z, the differences and the offsets from one trip to the next (f3 = U*d3, constant along a chain):
//   D=1: s[u] constant;  D=2: s[u] += u*(U*d2);  D=3: s[u] += u*E + C(u,2)*F, E = U*d2 + C(U,2)*d3, F = U*d3.
template <int D, u32 U> __device__ __forceinline__ void qcf_filter_advance(u32& z, u32& d1, u32& d2, const u32 d3, const u32 f3, u32 (&s)[U])
{
    const u32 e2 = U * d2 + (U * (U - 1u) / 2u) * d3;
    z += U * d1 + (U * (U - 1u) / 2u) * d2 + (U * (U - 1u) * (U - 2u) / 6u) * d3;
    d1 += e2;
    if (D >= 3) {
        d2 += f3;
    }
    if (D >= 2) {
#pragma unroll
        for (u32 u = 1; u < U; ++u) {
            s[u] += u * e2;
            if (D >= 3) {
                s[u] += (u * (u - 1u) / 2u) * f3;
            }
        }
    }
}

// ---- degree 1: the packed 16-bit trip ---------------------------------------------------------------------------------
// At degree 1 the trip offsets are constant inside a segment, s[u] = u*d1, and nothing is left for the multiply pipe to do,
// so the trip is bound by the integer ALU pipe (one fused add+min per guess at half rate).  The packed form halves that:
// two guesses share one VIADDMNMX.U16x2, which adds and compares the HIGH HALVES of the tracked limb only.
//   true value of step u:   v = z + s[u] (mod 2^32),   v >> 16 = (zh + sh + cy) mod 2^16,  zh = z >> 16, sh = s[u] >> 16,
//                           cy = carry out of the low halves, 0 or 1
//   packed value:           c = (zh + sh + 1) mod 2^16 = (v >> 16) + 1 - cy   in { v >> 16, (v >> 16) + 1 }
//   v <= bound  =>  v >> 16 <= bound >> 16  =>  c <= (bound >> 16) + 1 < T,   T = (bound >> 16) + 2   (no wrap: T <= 0xFFFF).
// So "some packed value of the trip is below T" is necessary for "some step of the trip is at or below the bound"; a trip
// that raises it is replayed with the full 32-bit comparison (rare: 32 * T / 2^16 per lane and trip).  The +1 lives in the
// packed offsets, so the trip needs one PRMT (broadcast of zh) and U/2 packed add+min.
__device__ __forceinline__ u32 qcf_addmin16x2(u32 a, u32 b, u32 c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmin_u16x2(a, b, c); // per half: min(a + b mod 2^16, c)
#else
    const u32 lo = (a + b) & 0xFFFFu, hi = ((a >> 16) + (b >> 16)) & 0xFFFFu;
    return ((lo < (c & 0xFFFFu)) ? lo : (c & 0xFFFFu)) | (((hi < (c >> 16)) ? hi : (c >> 16)) << 16);
#endif
}
// (x >> 16) | (y & 0xFFFF0000): the high halves of two values in one register
__device__ __forceinline__ u32 qcf_pack_hi16(u32 x, u32 y)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, 0x7632);
#else
    return (x >> 16) | (y & 0xFFFF0000u);
#endif
}
// packed offsets of a trip at degree 1: half h of S2[p] = ((2p + h) * d1 >> 16) + 1  (mod 2^16)
template <u32 U> __device__ __forceinline__ void qcf_filter_pack_offsets(u32 (&S2)[U / 2], u32 d1)
{
#pragma unroll
    for (u32 p = 0; p < U / 2; ++p) {
        S2[p] = qcf_pack_hi16((2u * p) * d1 + 0x10000u, (2u * p + 1u) * d1 + 0x10000u);
    }
}
// One packed trip: folds the U packed values of the trip starting at tracked limb z into the running minima m0, m1
// (both start at T | T << 16); returns true when some packed value was below T since the minima were last reset.
template <u32 U> __device__ __forceinline__ bool qcf_filter_trip16(u32 z, const u32 (&S2)[U / 2], u32& m0, u32& m1, u32 t2)
{
    const u32 zz = qcf_pack_hi16(z, z);
#pragma unroll
    for (u32 p = 0; p < U / 2; p += 2) {
        m0 = qcf_addmin16x2(zz, S2[p], m0);
        m1 = qcf_addmin16x2(zz, S2[p + 1], m1);
    }
    return ((m0 ^ t2) | (m1 ^ t2)) != 0u;
}
// Entering a segment at degree 1: as qcf_filter_enter_segment, without the trip offsets (the packed ones are rebuilt from d1).
__device__ __forceinline__ void qcf_filter_enter_segment1(u32& z, u32& d1, u32& elo, u32 shift, u32 nu_lo, u32 nu_hi, u32 rho32, u32 kappa)
{
    z += shift - (u32)(((u64)rho32 * kappa) >> 32);
    const u32 lo = elo + nu_lo;
    d1 += nu_hi + ((lo < elo) ? 1u : 0u);
    elo = lo;
}

// ---- degree 2: packed 16-bit sub-trips ----------------------------------------------------------------------------------
// At degree 2 the trip offsets move: s_j[u] = s_0[u] + j * u * e, e = (sub-trip length) * d2, which costs one multiply-add
// per guess and trip on top of the add+min.  When e is a multiple of 2^16 (the host checks it: d2 is a multiple of
// 2^(33 - t), t = the stride's 2-adic valuation) the HIGH HALVES move exactly too, hi16(s_j[u]) = hi16(s_0[u]) + j * (u*e >> 16)
// mod 2^16, and two guesses share a register as at degree 1.  A 32-bit multiply-add X = S2 + j * E2 on two packed halves lets
// the low half carry into the high one, cy_j times with 0 <= cy_j <= j.  So a GROUP of G sub-trips forms its offsets from the
// group's exact base S2 by ONE multiply-add per register (j = 1..G-1, on the other pipe) and the base itself moves by one
// exact packed add per group (EG = G * E2 per half):
//      per sub-trip of 16 guesses:  8 packed add+min,  8 multiply-adds (7 of 8 sub-trips);   per group: 8 packed adds
// -- 0.56 ALU-pipe instructions per guess instead of 1.  The thresholds absorb the roundings: low halves T = (bound >> 16) + 2
// as at degree 1, high halves T + G/2.
// The carries are kept within G/2 by ROUNDING the multiplier (setup only, nothing in the loop): with f = (low half of E2) / 2^16
// and s = (low half of S2) / 2^16 the carry is cy_j = floor(s + j*f); a register with f >= 1/2 takes one off the high half of
// its multiplier and adds G/2 to the high half of its base, so the high half of X is the true one plus
//     floor(s + j*f)            in [0, G/2]   (f <  1/2:  s + j*f < 1 + (G-1)/2),
//     floor(s - j*(1-f)) + G/2  in [0, G/2]   (f >= 1/2:  s - j*(1-f) > -(G-1)/2).
// (With plain multipliers the carries reach G - 1 and the high halves' threshold was T + 7: 5.8 units of 2^16 flagged per
// guess on average instead of 4.3; the warp-wide replays of flagged trips were a fifth of the stall samples at the top of a
// 96-bit range.)  A trip (two sub-trips) that raises a flag is replayed at full width.
#define QCF_FILTER_SUB 16u  // guesses per sub-trip
#define QCF_FILTER_GROUP 8u // sub-trips per group = 4 trips of 32
__device__ __forceinline__ u32 qcf_add16x2(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b); // per half, modulo 2^16
#else
    return ((a + b) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
#endif
}
// Packed offsets of a group starting where the first differences are d1, d2: low half of S2[p] = (s[2p] >> 16) + 1, high half
// (s[2p + 1] >> 16) + 1 (+ G/2 when the register's multiplier is rounded down, see above),  s[u] = u*d1 + C(u,2)*d2.
// E2 = the MULTIPLIER of the sub-trip index: the move per sub-trip (e = 16 * d2 is a multiple of 2^16), its high half one short
// when its low half is at least 2^15; EG = the exact move per group, per half modulo 2^16.
#define QCF_FILTER_GROUP_SLACK (QCF_FILTER_GROUP / 2u) // what the high halves' threshold is widened by
__device__ __forceinline__ u32 qcf_filter_pack2_rounded(u32 e2) // 1 when the multiplier's high half was rounded down
{
    return (e2 >> 15) & 1u;
}
__device__ __forceinline__ void qcf_filter_pack2_base(u32 (&S2)[QCF_FILTER_SUB / 2], const u32 (&E2)[QCF_FILTER_SUB / 2], u32 d1, u32 d2)
{
#pragma unroll
    for (u32 p = 0; p < QCF_FILTER_SUB / 2; ++p) {
        const u32 a = 2u * p, b = 2u * p + 1u;
        static_assert(QCF_FILTER_GROUP_SLACK == 4u, "the bias below is (bit 15 of the multiplier) << 3 = 4 << 16");
        S2[p] = qcf_pack_hi16(a * d1 + (a * (a - 1u) / 2u) * d2 + 0x10000u, b * d1 + (b * (b - 1u) / 2u) * d2 + 0x10000u)
            + ((E2[p] & 0x8000u) << 3);
    }
}
__device__ __forceinline__ void qcf_filter_pack2_moves(u32 (&E2)[QCF_FILTER_SUB / 2], u32 (&EG)[QCF_FILTER_SUB / 2], u32 d2)
{
    // e = 16 * d2 is a multiple of 2^16, so the halves of the move of register p are (2p * e >> 16, (2p + 1) * e >> 16): register 0
    // holds (0, e >> 16) = e itself and every next one is a packed add of (2e >> 16) in both halves further -- 2 packed adds per
    // register for both tables instead of 4 multiplies and 2 permutes
    const u32 e = QCF_FILTER_SUB * d2, eg = QCF_FILTER_GROUP * e;
    const u32 step = qcf_pack_hi16(2u * e, 2u * e), step_g = qcf_pack_hi16(2u * eg, 2u * eg);
    u32 m = e, g = eg;
#pragma unroll
    for (u32 p = 0; p < QCF_FILTER_SUB / 2; ++p) {
        E2[p] = m - ((m & 0x8000u) << 1); // rounded: the high half one short when the low half is at least 2^15
        EG[p] = g;
        m = qcf_add16x2(m, step);
        g = qcf_add16x2(g, step_g);
    }
}
// The base one sub-trip on, exactly (runs shorter than a group): the move is the multiplier with its rounding undone.
__device__ __forceinline__ void qcf_filter_pack2_step_base(u32 (&S2)[QCF_FILTER_SUB / 2], const u32 (&E2)[QCF_FILTER_SUB / 2])
{
#pragma unroll
    for (u32 p = 0; p < QCF_FILTER_SUB / 2; ++p) {
        S2[p] = qcf_add16x2(S2[p], E2[p] + (qcf_filter_pack2_rounded(E2[p]) << 16)); // the add is per half
    }
}
// Sub-trip j of a group (0 <= j < G): folds its 16 packed values into the running minima; z = the tracked limb at its first step.
__device__ __forceinline__ void qcf_filter_subtrip16(u32 z, u32 j, const u32 (&S2)[QCF_FILTER_SUB / 2], const u32 (&E2)[QCF_FILTER_SUB / 2], u32& m0, u32& m1)
{
    const u32 zz = qcf_pack_hi16(z, z);
#pragma unroll
    for (u32 p = 0; p < QCF_FILTER_SUB / 2; p += 2) {
        m0 = qcf_addmin16x2(zz, E2[p] * j + S2[p], m0);
        m1 = qcf_addmin16x2(zz, E2[p + 1] * j + S2[p + 1], m1);
    }
}
// The tracked limb one sub-trip on (exact).  The loop does not carry the first difference d1 itself but the limb's move over
// one sub-trip,  a = SUB*d1 + C(SUB,2)*d2,  which grows by SUB^2 * d2 per sub-trip: two additions per sub-trip instead of the
// seven instructions per trip that stepping z and d1 took.  Where d1 is needed -- entering a segment, replaying a flagged
// trip -- it is  d1z + k*d2  (k = the step index, d1z = the first difference extrapolated back to step 0, which only a
// segment entry changes).
__device__ __forceinline__ u32 qcf_filter_sub_move(u32 d1, u32 d2)
{
    return QCF_FILTER_SUB * d1 + (QCF_FILTER_SUB * (QCF_FILTER_SUB - 1u) / 2u) * d2;
}
__device__ __forceinline__ void qcf_filter_sub_advance(u32& z, u32& a, u32 d2)
{
    z += a;
    a += (QCF_FILTER_SUB * QCF_FILTER_SUB) * d2;
}
// thresholds of the packed test in both halves: low T, high T + G/2
__device__ __forceinline__ u32 qcf_filter_t2_group(u32 bound)
{
    const u32 t = (bound >> 16) + 2u;
    return ((t + QCF_FILTER_GROUP_SLACK) << 16) | t;
}

// ---- replay of a flagged packed trip ------------------------------------------------------------------------------------
// The running minima of a packed trip keep four classes apart: m0 folds the registers p = 0, 2, .., m1 the odd ones, and a
// register holds step 2p in its low half and 2p + 1 in its high half -- so low m0, high m0, low m1, high m1 are the steps
// u = 0, 1, 2, 3 (mod 4) of the trip.  A raised flag therefore names the residue of u mod 4 of every step that may be at or
// below the bound, and the full-width check of the trip only has to look at that quarter: U / 4 steps, evaluated directly
// (independent multiply-adds instead of a chain of U dependent additions).  The warp-wide full-width minimum over all U
// steps was 12 % of the issued instructions and 22 % of the stall samples of the 96-bit bench launches (ncu source page).
__device__ __forceinline__ u32 qcf_filter_flag_classes(u32 m0, u32 m1, u32 t2)
{
    const u32 x0 = m0 ^ t2, x1 = m1 ^ t2;
    return ((x0 & 0xFFFFu) ? 1u : 0u) | ((x0 >> 16) ? 2u : 0u) | ((x1 & 0xFFFFu) ? 4u : 0u) | ((x1 >> 16) ? 8u : 0u);
}
__device__ __forceinline__ u32 qcf_lowest_bit_index(u32 x) // x != 0
{
#if defined(__CUDA_ARCH__)
    return (u32)__ffs((int)x) - 1u;
#else
    return (u32)__builtin_ctz(x);
#endif
}
// min over the steps u = r, r + 4, .., r + U - 4 of a trip of  v(u) = z + u*d1 + C(u,2)*d2  (d2 = 0 at degree 1):
// along the residue class the value is  A + i*D1 + C(i,2)*D2  with  A = v(r),  D1 = 4*d1 + (4r + 6)*d2,  D2 = 16*d2.
__device__ __forceinline__ void qcf_filter_class_start(u32 z, u32 d1, u32 d2, u32 r, u32& a, u32& e1, u32& e2)
{
    a = z + r * d1 + ((r * (r - 1u)) >> 1) * d2;
    e1 = 4u * d1 + (4u * r + 6u) * d2;
    e2 = 16u * d2;
}
template <u32 U> __device__ __forceinline__ u32 qcf_filter_class_min(u32 z, u32 d1, u32 d2, u32 r)
{
    u32 a, e1, e2;
    qcf_filter_class_start(z, d1, d2, r, a, e1, e2);
    u32 mn = a;
#pragma unroll
    for (u32 i = 1; i < U / 4u; ++i) {
        const u32 v = a + i * e1 + (i * (i - 1u) / 2u) * e2;
        mn = (v < mn) ? v : mn;
    }
    return mn;
}
// The full-width minimum of a flagged trip over the classes `cls` names, for the whole warp together: one round per class
// some lane has flagged (almost always one round; a lane without that class computes along and ignores the result).
// 0xFFFFFFFF for a lane without flags.
template <u32 U> __device__ __forceinline__ u32 qcf_filter_flagged_min(u32 z, u32 d1, u32 d2, u32 cls)
{
    u32 mn = 0xFFFFFFFFu;
#if defined(__CUDA_ARCH__)
    do {
        const u32 mr = qcf_filter_class_min<U>(z, d1, d2, cls ? qcf_lowest_bit_index(cls) : 0u);
        mn = (cls && (mr < mn)) ? mr : mn;
        cls &= cls - 1u;
    } while (__any_sync(0xffffffffu, cls != 0u));
#else
    for (; cls; cls &= cls - 1u) {
        const u32 mr = qcf_filter_class_min<U>(z, d1, d2, qcf_lowest_bit_index(cls));
        mn = (mr < mn) ? mr : mn;
    }
#endif
    return mn;
}
