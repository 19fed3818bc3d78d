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
    if (LG == 1) {
        a[0] += b[0];
    } else if (LG == 2) {
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[0]), "+r"(a[LG - 1]) : "r"(b[0]), "r"(b[LG - 1]));
    } else if (LG == 3) {
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
            : "+r"(a[0]), "+r"(a[1]), "+r"(a[LG - 1])
            : "r"(b[0]), "r"(b[1]), "r"(b[LG - 1]));
    } else {
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
