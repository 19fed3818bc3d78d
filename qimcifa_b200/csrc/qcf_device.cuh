// Device-side limb arithmetic for the sweep kernels (sm_100a).
//
// Replaces the generic BigInteger operator% of the reference's hot line
//     if ((toFactor % base) == 0U)            reference include/qimcifa.hpp:369
// (boost::multiprecision::cpp_int there; include/big_integer.hpp / src/common/big_integer.cpp:136-228
// in the pure-language backend) with a fixed-width exact divisibility test on 32-bit limbs.
#pragma once

#include <cstdint>

typedef uint32_t u32;
typedef uint64_t u64;

// The arithmetic below is plain integer C++ apart from the carry-flag adds, so a host compiler can build it too:
// tests/device_math_host.cpp includes this header with g++ and checks every function against Python integers
// (tests/test_device_math.py) -- the same source the kernels are compiled from, not a restatement.
#if !defined(__CUDACC__)
#define __device__
#define __forceinline__ inline
#endif

// a*b + c as a 64-bit value through the 32x32+64 multiply-add (IMAD.WIDE.U32, full rate on the B200), and the high word
// of such a value.  The plain C++ form of "high word of a*b + c" compiles to IMAD.HI.U32, which the B200 issues at 0.4 of
// the IMAD rate (qcf_microbench kind 3: 7.4e12 against 1.86e13 per second) -- on the pipe that bounds the exact kernels.
__device__ __forceinline__ u64 qcf_madwide(u32 a, u32 b, u64 c)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return (u64)a * b + c;
#endif
}
__device__ __forceinline__ u32 qcf_hi32(u64 x)
{
#if defined(__CUDA_ARCH__)
    u32 lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
    (void)lo;
    return hi;
#else
    return (u32)(x >> 32);
#endif
}

// g^-1 mod 2^32 for odd g: (3g)^2 is correct to 5 bits, three Newton steps x <- x(2 - gx) give 40.
__device__ __forceinline__ u32 qcf_inv32(u32 g)
{
    u32 x = (3u * g) ^ 2u;
    x *= 2u - g * x;
    x *= 2u - g * x;
    x *= 2u - g * x;
    return x;
}

// Exact test "g divides N" for odd g (every wheel guess is odd), by word-serial Montgomery/Hensel
// reduction: QN times, q = t0 * (-g^-1) mod 2^32, t = (t + q*g) / 2^32.  After QN words
//     t = (N + M*g) / 2^(32 QN),   M = -N/g mod 2^(32 QN),
// so if g | N with cofactor C = N/g < 2^(32 QN) then M = 2^(32 QN) - C and t == g exactly; conversely
// t == g forces N = (2^(32 QN) - M) * g.  The caller guarantees N / g < 2^(32 QN) for every guess of
// the launch (QN is chosen per launch from the smallest guess), so "t == g" is necessary and sufficient.
//   LN = limbs of N, LG = limbs of g, QN = quotient words.
template <int LN, int LG, int QN> __device__ __forceinline__ bool qcf_divides(const u32 (&n)[LN], const u32 (&g)[LG])
{
    constexpr int LT = ((LN > QN + LG) ? LN : (QN + LG)) + 1;
    u32 t[LT];
#pragma unroll
    for (int i = 0; i < LT; ++i) {
        t[i] = (i < LN) ? n[i] : 0u;
    }
    const u32 ninv = 0u - qcf_inv32(g[0]);
#pragma unroll
    for (int s = 0; s < QN; ++s) {
        const u32 q = t[s] * ninv;
        u64 c = 0;
#pragma unroll
        for (int k = 0; k < LG; ++k) {
            const u64 acc = (u64)q * g[k] + t[s + k] + c;
            t[s + k] = (u32)acc;
            c = acc >> 32;
        }
#pragma unroll
        for (int k = s + LG; k < LT; ++k) {
            const u64 acc = (u64)t[k] + c;
            t[k] = (u32)acc;
            c = acc >> 32;
        }
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < LG; ++k) {
        ok = ok && (t[QN + k] == g[k]);
    }
#pragma unroll
    for (int k = QN + LG; k < LT; ++k) {
        ok = ok && (t[k] == 0u);
    }
    return ok;
}

// r = a + b (LG limbs + 32-bit), wrap-around at LG limbs.
template <int LG> __device__ __forceinline__ void qcf_add32(u32 (&r)[LG], const u32 (&a)[LG], u32 b)
{
    u64 c = b;
#pragma unroll
    for (int k = 0; k < LG; ++k) {
        c += a[k];
        r[k] = (u32)c;
        c >>= 32;
    }
}

// a += b (LG limbs): one add per limb on the hardware carry flag (the 64-bit-accumulator form compiles to
// compare + conditional increment, two instructions more per limb)
template <int LG> __device__ __forceinline__ void qcf_addn(u32 (&a)[LG], const u32 (&b)[LG])
{
#if defined(__CUDA_ARCH__)
    if (LG == 1) {
        a[0] += b[0];
    } else if (LG == 2) {
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[0]), "+r"(a[LG - 1]) : "r"(b[0]), "r"(b[LG - 1]));
    } else if (LG == 3) {
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
            : "+r"(a[0]), "+r"(a[1]), "+r"(a[LG - 1])
            : "r"(b[0]), "r"(b[1]), "r"(b[LG - 1]));
    } else
#endif
    {
        u64 c = 0;
#pragma unroll
        for (int k = 0; k < LG; ++k) {
            c += (u64)a[k] + b[k];
            a[k] = (u32)c;
            c >>= 32;
        }
    }
}

// a < b
template <int LG> __device__ __forceinline__ bool qcf_less(const u32 (&a)[LG], const u32 (&b)[LG])
{
    bool lt = false;
#pragma unroll
    for (int k = 0; k < LG; ++k) { // low to high: the highest differing limb decides
        lt = (a[k] < b[k]) || ((a[k] == b[k]) && lt);
    }
    return lt;
}

// ---- reductions with the reciprocal supplied by the caller (chain kernels: it is carried from guess to guess) ----

// t == g after QN words of  t = (t + q*g)/2^32,  q = t0 * ninv  (see qcf_divides<> in qcf_device.cuh for the proof)
template <int LN, int LG, int QN> __device__ __forceinline__ bool qcf_divides_ninv(const u32 (&n)[LN], const u32 (&g)[LG], u32 ninv)
{
    constexpr int LT = ((LN > QN + LG) ? LN : (QN + LG)) + 1;
    u32 t[LT];
#pragma unroll
    for (int i = 0; i < LT; ++i) {
        t[i] = (i < LN) ? n[i] : 0u;
    }
#pragma unroll
    for (int s = 0; s < QN; ++s) {
        const u32 q = t[s] * ninv;
        u64 c = 0;
#pragma unroll
        for (int k = 0; k < LG; ++k) {
            const u64 acc = (u64)q * g[k] + t[s + k] + c;
            t[s + k] = (u32)acc;
            c = acc >> 32;
        }
#pragma unroll
        for (int k = s + LG; k < LT; ++k) {
            const u64 acc = (u64)t[k] + c;
            t[k] = (u32)acc;
            c = acc >> 32;
        }
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < LG; ++k) {
        ok = ok && (t[QN + k] == g[k]);
    }
#pragma unroll
    for (int k = QN + LG; k < LT; ++k) {
        ok = ok && (t[k] == 0u);
    }
    return ok;
}

// Two-limb guesses (every sweep of a 65..128-bit semiprime): the same reduction written so that each quotient word is
// two 32x32+64 multiply-adds (IMAD.WIDE) and two 64-bit adds.
template <int LN, int QN> __device__ __forceinline__ bool qcf_divides_ninv2(const u32 (&n)[LN], const u32 g0, const u32 g1, u32 ninv)
{
    constexpr int LT = ((LN > QN + 2) ? LN : (QN + 2)) + 2;
    u32 t[LT];
#pragma unroll
    for (int i = 0; i < LT; ++i) {
        t[i] = (i < LN) ? n[i] : 0u;
    }
#pragma unroll
    for (int s = 0; s < QN; ++s) {
        const u32 q = t[s] * ninv;
        const u64 a = (u64)q * g0 + t[s];                              // low word becomes 0
        const u64 b = (u64)q * g1 + t[s + 1] + (a >> 32);              // < 2^64: (2^32-1)^2 + 2(2^32-1)
        t[s + 1] = (u32)b;
        const u64 c = (((u64)t[s + 3] << 32) | t[s + 2]) + (b >> 32);  // carry into the two limbs above
        t[s + 2] = (u32)c;
        t[s + 3] = (u32)(c >> 32);
        if (s + 4 < LT) {
            t[s + 4] += (c < (b >> 32)) ? 1u : 0u;
        }
    }
    bool ok = (t[QN] == g0) && (t[QN + 1] == g1);
#pragma unroll
    for (int k = QN + 2; k < LT; ++k) {
        ok = ok && (t[k] == 0u);
    }
    return ok;
}

// The common shape -- two-limb guess, two quotient words, N of 3 or 4 limbs (96..128-bit semiprimes near sqrt(N)) --
// with the carries kept in 64-bit accumulators: per word one low multiply, two 32x32+64 multiply-adds, two 64-bit adds.
template <int LN> __device__ __forceinline__ bool qcf_divides_q2g2(const u32 (&n)[LN], const u32 g0, const u32 g1, u32 ninv)
{
    const u32 n3 = (LN > 3) ? n[LN - 1] : 0u;
    const u32 q0 = n[0] * ninv;
    const u64 a0 = (u64)q0 * g0 + n[0];                  // low word 0
    const u64 b0 = (u64)q0 * g1 + n[1] + (a0 >> 32);     // new limb 1 (low), carry (high)
    // limbs 2,3.  For a true divisor N + M*g == g*2^64 < 2^128, so no partial sum can wrap; a wrap (possible only for
    // a non-divisor of a 4-limb N) subtracts 2^128 from the total, and N = (2^64 - M)*g + 2^128 is impossible for N < 2^128:
    // the wrapped value can never equal g.
    const u64 t0 = (((u64)n3 << 32) | n[2]) + (b0 >> 32);
    const u32 q1 = (u32)b0 * ninv;
    const u64 a1 = (u64)q1 * g0 + (u32)b0;
    const u64 b1 = (u64)q1 * g1 + (u32)t0 + (a1 >> 32);  // final limb 0 (low), carry (high)
    const u64 t1 = (t0 >> 32) + (b1 >> 32);              // final limb 1 (+ overflow bit): must equal g1
    return ((u32)b1 == g0) && (t1 == (u64)g1);
}

// ---- two-stage exact test (chain kernel, two-limb guess, two quotient words) ----
//
// One Newton step on the NEGATED reciprocal y = -g^-1 mod 2^32 (what the reduction multiplies by): with x = -y,
// x*(2 - g*x) = -y*(2 + g*y), so y' = y*(2 + g*y) -- two multiply-adds and no negation.  As for x, the step is exact
// when y is already correct to 16 bits for the new g (2^16 | g' - g).
__device__ __forceinline__ u32 qcf_ninv_step(u32 y, u32 g0)
{
    return y * (g0 * y + 2u);
}

// Stage 1: limb 0 of the final t of qcf_divides_q2g2 and nothing else.  W = 2^32, N = n0 + W*N1, g = g0 + W*g1.
//   T1 = (N + q0*g)/W = N1 + h0 + q0*g1,          h0 = (n0 + q0*g0)/W  (exact: q0 makes the low word vanish)
//   T2 = (T1 + q1*g)/W = floor(T1/W) + h1 + q1*g1, h1 = (L1 + q1*g0)/W, L1 = T1 mod W
// so T2 mod W depends on T1 mod W^2 only:  t1 = (n1 + n2*W + h0 + q0*g1) mod W^2 (n3 and every carry past bit 64 drop
// out), the high word of (q1*g0 + t1) mod W^2 is (h1 + floor(t1/W)) mod W, and adding the low word of q1*g1 gives
// T2 mod W.  g | N  =>  T2 == g  =>  T2 mod W == g0: a true divisor always passes; a non-divisor passes with
// probability 2^-32 and is then decided by the full test (stage 2), so the result is that of the full test.
// Five multiply instructions (IMAD.HI, IMAD.WIDE, IMAD, IMAD.HI, IMAD; q0 comes from the caller), two adds and a compare: no
// carry out of bit 64 is formed and no 64-bit operand is assembled from halves.
// q0 = n0 * ninv comes from the caller: along a chain it is carried by one add per guess (qcf_ninv_delta below).
// Plain C++: ptxas turns each "high word of a*b + c" into ONE IMAD.HI.U32 with the 64-bit addend folded in.  Spelling the
// products as IMAD.WIDE with explicit carries was measured on the B200 and is slower (profiles/r02_exact_chain_variants.txt:
// 1.37e12 against 1.45e12 guesses/s): a 64-bit IMAD.WIDE result occupies the multiplier pipe as long as an IMAD.HI, and the
// carries cost issue slots.
template <int LN> __device__ __forceinline__ bool qcf_lowlimb_q2g2_q0(const u32 (&n)[LN], const u32 q0, const u32 g0, const u32 g1, const u32 ninv)
{
    const u64 n12 = ((u64)n[2] << 32) | n[1];
    const u32 h0 = (u32)(((u64)q0 * g0 + n[0]) >> 32);
    const u64 t1 = (u64)q0 * g1 + n12 + h0; // mod W^2: unsigned wrap-around is the intended arithmetic
    const u32 q1 = (u32)t1 * ninv;
    const u64 a1 = (u64)q1 * g0 + t1;       // low word 0
    const u32 t2 = (u32)(a1 >> 32) + q1 * g1;
    return t2 == g0;
}
template <int LN> __device__ __forceinline__ bool qcf_lowlimb_q2g2(const u32 (&n)[LN], const u32 g0, const u32 g1, const u32 ninv)
{
    return qcf_lowlimb_q2g2_q0<LN>(n, n[0] * ninv, g0, g1, ninv);
}

// Along a chain g_k = g_0 + k*S with 2^16 | S the reciprocal is LINEAR in k modulo 2^32:  g_k^-1 = x0 * (1 - k*S*x0 + (k*S*x0)^2 - ...)
// and every term from the square on vanishes, so with y = -g^-1 (what the reductions multiply by)
//     y_k = y_0 + k * dy,   dy = S * y_0^2 mod 2^32,
// and the first quotient word  q0_k = n0 * y_k = q0_0 + k * (n0 * dy)  is linear too: one add each per guess, on the ALU
// pipe, instead of a Newton step (two multiplies) and the q0 multiply on the pipe that bounds the kernel.
__device__ __forceinline__ u32 qcf_ninv_delta(u32 y0, u32 stride_lo) { return stride_lo * y0 * y0; }

// Stage 1 for any shape: limb 0 of the final t of qcf_divides<LN, LG, QN> (t[QN] in its notation) and nothing above it.
// The word-serial reduction is run on limbs 0..QN only -- limb QN itself modulo 2^32 -- because no limb above QN feeds a
// lower one.  g | N  =>  final t == g  =>  t[QN] == g[0]; a non-divisor passes with probability 2^-32 and is decided by
// the full test, so "stage 1 and then the full test" gives the result of the full test.  ninv = -g^-1 mod 2^32.
template <int LN, int LG, int QN> __device__ __forceinline__ bool qcf_lowlimb(const u32 (&n)[LN], const u32 (&g)[LG], const u32 ninv)
{
    u32 t[QN + 1];
#pragma unroll
    for (int i = 0; i <= QN; ++i) {
        t[i] = (i < LN) ? n[i] : 0u;
    }
#pragma unroll
    for (int s = 0; s < QN; ++s) {
        const u32 q = t[s] * ninv;
        u64 c = ((u64)q * g[0] + t[s]) >> 32; // the low word vanishes
#pragma unroll
        for (int k = 1; k < LG; ++k) {
            if (s + k <= QN) {
                const u64 acc = (u64)q * g[k] + t[s + k] + c;
                t[s + k] = (u32)acc;
                c = acc >> 32;
            }
        }
#pragma unroll
        for (int k = s + LG; k <= QN; ++k) {
            const u64 acc = (u64)t[k] + c;
            t[k] = (u32)acc;
            c = acc >> 32;
        }
    }
    return t[QN] == g[0];
}
