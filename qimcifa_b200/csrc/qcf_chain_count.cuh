// Helpers shared by the chain-form kernels (filter, exact chain, gcd chain) for the wheel primes of the requested level
// that the chain's tables do not remove ("extra" primes: superset mode, and prime 29 at level 10, whose period
// 2*3*...*29 does not fit the 32-bit residue tables).  Used on rare paths only: reporting a hit and the device-side
// count of QCF_COUNT_ON_DEVICE.
#pragma once

#include "qcf_device.cuh"
#include "qcf_internal.h"

// g mod q for a small prime q
template <int LG> __device__ __forceinline__ u32 qcf_mod_small(const u32 (&g)[LG], u32 q)
{
    const u32 r32 = (u32)((1ull << 32) % q);
    u32 r = 0;
#pragma unroll
    for (int i = LG - 1; i >= 0; --i) {
        r = (u32)(((u64)r * r32 + g[i] % q) % q);
    }
    return r;
}

// true when g is a multiple of one of the launch's extra primes: such a guess is not part of the level
template <int LG> __device__ __forceinline__ bool qcf_hits_extra(const FilterParams& prm, const u32 (&g)[LG])
{
    for (u32 e = 0; e < prm.n_extra; ++e) {
        if (qcf_mod_small<LG>(g, prm.extra[e]) == 0u) {
            return true;
        }
    }
    return false;
}

// Guesses of chain (c, g0s) coprime to the extra primes.
__device__ __noinline__ static u32 qcf_chain_count_extra(const FilterParams& prm, u64 c, u32 g0s)
{
    u32 res[QCF_MAX_EXTRA], inc[QCF_MAX_EXTRA];
    for (u32 e = 0; e < prm.n_extra; ++e) {
        const u32 q = prm.extra[e];
        // g0 = base + c*P + g0s (mod q);  stride = P << tprime (mod q)
        u32 b = 0;
        for (int i = 2; i >= 0; --i) {
            b = (u32)(((u64)b * ((1ull << 32) % q) + prm.base[i] % q) % q);
        }
        res[e] = (u32)((b + (c % q) * (prm.period % q) + g0s) % q);
        u32 st = prm.period % q;
        for (u32 k = 0; k < prm.tprime; ++k) {
            st = (st * 2u) % q;
        }
        inc[e] = st;
    }
    u32 cnt = 0;
    for (u32 k = 0; k < prm.steps; ++k) {
        bool ok = true;
        for (u32 e = 0; e < prm.n_extra; ++e) {
            ok = ok && (res[e] != 0u);
            res[e] += inc[e];
            res[e] = (res[e] >= prm.extra[e]) ? (res[e] - prm.extra[e]) : res[e];
        }
        cnt += ok ? 1u : 0u;
    }
    return cnt;
}
