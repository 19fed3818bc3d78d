// Chain arithmetic of the filter kernel (qcf_filter.cu), kept apart from the kernel so that a host compiler can build
// it too: tests/device_math_host.cpp walks whole chains with THIS code on the plans of the real planner and
// tests/test_filter_math.py checks every step against the definition in Python integers.  See qcf_filter.cu for the
// mathematics; everything here is exact arithmetic modulo 2^64 / 2^32.
#pragma once

#include "qcf_device.cuh"

// guesses per trip: 32 at degree <= 2, 16 at degree 3
#ifndef QCF_FILTER_UNROLL_FOR
#define QCF_FILTER_UNROLL_FOR(D) (((D) >= 3) ? 16 : 32)
#endif

// Chain setup for the chain starting at guess g0 (only g0 mod 2^64 matters): x = g0^-1 mod 2^64, then the tracked high
// limb z of Z(g_0) (started `slack` too high, see qcf_filter.cu) and the forward differences d1..d3 of Z's high limb.
//   n64 = N mod 2^64, P = period, tp = log2 of the stride in periods, al = alignment shift, c_lo = cofactor floor.
template <int D> __device__ __forceinline__ void qcf_filter_chain_init(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u32 slack, u32& z, u32& d1, u32& d2, u32& d3)
{
    u64 x = qcf_inv32((u32)g0);
    x *= 2ull - g0 * x;
    const u64 nx = n64 * x;
    const u64 y = ((u64)P * x) << tp; // S * x
    z = (u32)(((nx - c_lo) << al) >> 32) + slack;
    d1 = 0, d2 = 0, d3 = 0;
    const u64 z1 = 0ull - (nx << al) * y;
    if (D == 1) {
        d1 = (u32)(z1 >> 32);
    } else {
        const u64 z2 = 0ull - z1 * y;
        if (D == 2) {
            d1 = (u32)((z1 + z2) >> 32);
            d2 = (u32)((z2 + z2) >> 32);
        } else {
            const u64 z3 = 0ull - z2 * y;
            d1 = (u32)((z1 + z2 + z3) >> 32);
            d2 = (u32)((2ull * z2 + 6ull * z3) >> 32);
            d3 = (u32)((6ull * z3) >> 32);
        }
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
