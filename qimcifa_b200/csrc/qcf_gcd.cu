// Common-factor mode: the reference built with IS_RSA_SEMIPRIME=OFF reports a guess when gcd(guess, N) != 1
// (reference include/qimcifa.hpp:201-210 gcd(), :373-379 "Has common factor").
//
// A per-guess Euclid is ~1000 instructions on a GPU (data-dependent loop, divergent lanes).  The chain kernel here
// does what Pollard-rho implementations do instead: every lane multiplies the guesses of its chain into one
// Montgomery accumulator modulo N (one LN x LG limb product + LG reduction words per guess, no branch), and takes
// ONE gcd(accumulator, N) per block of guesses.  gcd(prod g_k, N) != 1 exactly when some g_k shares a factor with N
// (2^-32 is a unit modulo odd N, so the Montgomery factors do not matter); only then is the block replayed with the
// per-guess test below, which finds the guesses themselves.  The host strips the factors of two from N first (every
// guess is odd) and maps the reported guess to gcd(guess, N) as the reference does.
#include <cstdio>
#include <cstdlib>

#include "qcf_chain_count.cuh"
#include "qcf_device.cuh"
#include "qcf_internal.h"

typedef unsigned __int128 du128;

// gcd(a, b) != 1 for odd a > 0 (binary gcd on one 128-bit register pair; b may be 0 or even).
__device__ __forceinline__ bool qcf_gcd_not_one_64(u64 a, u64 b)
{
    if (b == 0ull) {
        return a != 1ull;
    }
    b >>= __ffsll((long long)b) - 1;
    while (a != b) {
        if (a > b) {
            const u64 t = a;
            a = b;
            b = t;
        }
        b -= a;
        b >>= __ffsll((long long)b) - 1;
    }
    return a != 1ull;
}

__device__ __noinline__ bool qcf_gcd_not_one_128(du128 a, du128 b)
{
    if (b == 0) {
        return a != 1;
    }
    while (((u64)b & 1ull) == 0ull) {
        const u64 lo = (u64)b;
        b >>= lo ? (__ffsll((long long)lo) - 1) : 64;
    }
    while (a != b) {
        if (a > b) {
            const du128 t = a;
            a = b;
            b = t;
        }
        b -= a;
        while (((u64)b & 1ull) == 0ull) {
            const u64 lo = (u64)b;
            b >>= lo ? (__ffsll((long long)lo) - 1) : 64;
        }
    }
    return a != 1;
}

// gcd(g, N) != 1 for one odd guess g > 1: Hensel-reduce N by g (t = N * 2^(-32 LN) mod g up to one multiple of g:
// 0 <= t <= g, and gcd(g, t) = gcd(g, N) because 2 is a unit modulo g), then a binary gcd of two values below 2^(32 LG).
template <int LN, int LG> __device__ __forceinline__ bool qcf_common_factor(const u32 (&n)[LN], const u32 (&g)[LG])
{
    constexpr int LT = LN + LG + 1;
    u32 t[LT];
#pragma unroll
    for (int i = 0; i < LT; ++i) {
        t[i] = (i < LN) ? n[i] : 0u;
    }
    const u32 ninv = 0u - qcf_inv32(g[0]);
#pragma unroll
    for (int s = 0; s < LN; ++s) {
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
    // t[LN .. LN+LG] holds the reduced value (<= g, so it fits LG limbs + possibly equality with g)
    if (LG <= 2) {
        const u64 gv = (LG == 2) ? (((u64)g[LG - 1] << 32) | g[0]) : (u64)g[0];
        const u64 tv = (LG == 2) ? (((u64)t[LN + LG - 1] << 32) | t[LN]) : (u64)t[LN];
        const u64 r = (t[LN + LG] != 0u || tv >= gv) ? (tv - gv) : tv; // t == g (or the 2^(32 LG) wrap of it) -> 0
        return qcf_gcd_not_one_64(gv, r);
    } else {
        du128 gv = 0, tv = 0;
#pragma unroll
        for (int k = LG - 1; k >= 0; --k) {
            gv = (gv << 32) | g[k];
        }
#pragma unroll
        for (int k = LG; k >= 0; --k) {
            tv = (tv << 32) | t[LN + k];
        }
        return qcf_gcd_not_one_128(gv, (tv >= gv) ? (tv - gv) : tv);
    }
}

template <int LG> __device__ __noinline__ void qcf_gcd_report(QcfHits* hits, const u32 (&g)[LG])
{
    const u32 slot = qcf_hit_slot(hits);
    if (slot < QCF_MAX_HITS) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            hits->g[slot][q] = (q < LG) ? g[q] : 0u;
        }
    }
}

// ---- row form: one per-guess test per thread and row (ragged ends, short intervals) -------------------------------
template <int LN, int LG> __global__ void __launch_bounds__(QCF_BLOCK) qcf_gcd_rows_kernel(const ExactParams prm, const u32 count)
{
    extern __shared__ u32 s_tab[];
    u32* s_a = s_tab;
    u32* s_b = s_tab + prm.n_a;
    for (u32 i = threadIdx.x; i < prm.n_a; i += blockDim.x) {
        s_a[i] = prm.d_a[i];
    }
    for (u32 i = threadIdx.x; i < prm.n_b; i += blockDim.x) {
        s_b[i] = prm.d_b[i];
    }
    __syncthreads();
    u32 n[LN], v_lo[LG], v_hi[LG];
#pragma unroll
    for (int k = 0; k < LN; ++k) {
        n[k] = prm.n[k];
    }
#pragma unroll
    for (int k = 0; k < LG; ++k) {
        v_lo[k] = prm.v_lo[k];
        v_hi[k] = prm.v_hi[k];
    }
    const u32 P = prm.period, n_a = prm.n_a, n_b = prm.n_b;
    u64 tested = 0;
    u32 iter = 0;
    for (u64 row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
        const u64 moff = row / n_b;
        const u32 bj = s_b[(u32)(row - moff * n_b)];
        // base = prm.base + moff * P
        u32 base[LG];
        {
            const u64 lo = (u64)(u32)moff * P;
            const u64 hi = (u64)(u32)(moff >> 32) * P + (lo >> 32);
            const u32 prod[3] = { (u32)lo, (u32)hi, (u32)(hi >> 32) };
            u64 c = 0;
#pragma unroll
            for (int k = 0; k < LG; ++k) {
                c += (u64)prm.base[k] + prod[k];
                base[k] = (u32)c;
                c >>= 32;
            }
        }
        for (u32 i = threadIdx.x; i < n_a; i += blockDim.x) {
            u32 g0 = s_a[i] + bj;
            g0 = (g0 >= P) ? (g0 - P) : g0;
            u32 g[LG];
            qcf_add32<LG>(g, base, g0);
            const bool in = !qcf_less<LG>(g, v_lo) && !qcf_less<LG>(v_hi, g);
            if (in && !(prm.extra_prime && (qcf_mod_small<LG>(g, prm.extra_prime) == 0u))) {
                ++tested;
                if (qcf_common_factor<LN, LG>(n, g)) {
                    qcf_gcd_report<LG>(prm.hits, g);
                }
            }
        }
        if (((++iter) & 15u) == 0u) {
            if (*(volatile u32*)&prm.hits->stop) {
                break;
            }
        }
    }
    if (count) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            tested += __shfl_down_sync(0xffffffffu, tested, off);
        }
        if ((threadIdx.x & 31u) == 0u) {
            atomicAdd((unsigned long long*)&prm.hits->tested, (unsigned long long)tested);
        }
    }
}

// ---- chain form: Montgomery product of the chain's guesses, one gcd per block of guesses --------------------------
#define QCF_GCD_BLOCK 128
#define QCF_GCD_ROWS 4
// guesses multiplied together between two gcds.  A binary gcd of two LN-limb values is ~5000 instructions of its warp (a
// data-dependent loop: the warp runs as long as its slowest lane), and there are two accumulators per thread: at 1024 guesses per
// gcd that was ~10 instructions per guess beside a product that now takes 27; the host makes the chains of this mode up to 8192
// steps long (launch_exact_values, qcf_api.cu) and a chain is one batch.
#define QCF_GCD_BATCH 8192

// acc <- acc * g * 2^(-32 LG) mod N, kept below N.  acc < N < 2^(32 LN), g < 2^(32 LG), N odd, ninv = -N^-1 mod 2^32.
// Montgomery product by columns (finely integrated product scanning): ONE wide running sum S takes every partial product of
// a column -- acc[i] * g[k-i] and q[j] * n[k-j] -- emits the column's limb and moves on, S >>= 32.  Written with a 128-bit S,
// ptxas makes each term one IMAD.WIDE.U32 whose 64-bit addend is the running sum itself, carry-out in a predicate, and joins the
// carries with three-input IADD3.X: 26 instructions for the 12 multiply-adds of a 96-bit N and a two-limb guess.  The row form
// this replaces (round 2: one product row and one reduction row per limb of g, each row LN multiply-adds on the old limbs plus
// one carry chain) cost 67 in the loop: ptxas splits a mad.wide.u32 whose addend is a zero-extended 32-bit limb into multiply
// + add + add-with-carry, three instructions per term.
// LAZY: no final subtraction.  The result is t < acc * g / 2^(32 LG) + N; when every guess of the launch is below 2^gb and the top
// 32 LG - gb bits of N (as an LN-limb number) are not all ones, acc < 2^(32 LN) implies t < 2^(32 LN - (32 LG - gb)) + N <= 2^(32 LN):
// the accumulator stays within LN limbs without ever being reduced below N, and gcd(acc, N) does not care (the host checks the
// condition per launch, qcf_api.cu).
template <int LN, int LG, bool LAZY> __device__ __forceinline__ void qcf_montmul_acc(u32 (&acc)[LN], const u32 (&g)[LG], const u32 (&n)[LN], u32 ninv)
{
    u32 q[LG], t[LN + 1];
    du128 S = 0;
#pragma unroll
    for (int k = 0; k < LN + LG; ++k) {
#pragma unroll
        for (int i = 0; i < LN; ++i) {
            if ((k - i >= 0) && (k - i < LG)) {
                S += (u64)acc[i] * g[k - i];
            }
        }
#pragma unroll
        for (int j = 0; j < LG; ++j) {
            if ((j < k) && (k - j < LN)) {
                S += (u64)q[j] * n[k - j];
            }
        }
        if (k < LG) {
            q[k] = (u32)S * ninv; // makes the column's limb zero
            S += (u64)q[k] * n[0];
        } else {
            t[k - LG] = (u32)S;
        }
        S >>= 32;
    }
    t[LN] = (u32)S; // 0 or 1: t < acc + N < 2N
    if (LAZY) {
#pragma unroll
        for (int j = 0; j < LN; ++j) {
            acc[j] = t[j];
        }
        return;
    }
    // one conditional subtraction
    u32 d[LN];
    u64 b = 0;
#pragma unroll
    for (int j = 0; j < LN; ++j) {
        const u64 p = (u64)t[j] - n[j] - b;
        d[j] = (u32)p;
        b = (p >> 32) & 1ull;
    }
    const bool ge = (t[LN] != 0u) || (b == 0ull);
#pragma unroll
    for (int j = 0; j < LN; ++j) {
        acc[j] = ge ? d[j] : t[j];
    }
}

template <int LN, int LG, bool LAZY> __global__ void __launch_bounds__(QCF_GCD_BLOCK, 8) qcf_gcd_chain_kernel(const __grid_constant__ FilterParams prm)
{
    extern __shared__ u32 s_tab[];
    u32* s_a = s_tab;
    u32* s_b = s_tab + prm.n_a;
    for (u32 i = threadIdx.x; i < prm.n_a; i += blockDim.x) {
        s_a[i] = prm.d_a[i];
    }
    for (u32 i = threadIdx.x; i < prm.n_b; i += blockDim.x) {
        s_b[i] = prm.d_b[i];
    }
    __shared__ u64 s_claim;
    const u32 P = prm.period, n_a = prm.n_a, n_b = prm.n_b, kc = prm.steps;
    const u64 n_srows = (prm.n_rows + QCF_GCD_ROWS - 1u) / QCF_GCD_ROWS;
    u32 n[LN], stride[LG];
#pragma unroll
    for (int i = 0; i < LN; ++i) {
        // N in vector registers, not in the constant bank: a multiply-add whose factor is a uniform operand cannot take the running
        // 64-bit sum as its addend with a carry-out (ptxas then emits multiply + add + add-with-carry: 126 instead of 96 instructions
        // per two guesses in the loop)
#if defined(__CUDA_ARCH__)
        asm volatile("mov.b32 %0, %1;" : "=r"(n[i]) : "r"(prm.n[i]));
#else
        n[i] = prm.n[i];
#endif
    }
#pragma unroll
    for (int i = 0; i < LG; ++i) {
        stride[i] = prm.stride[i];
    }
    const u32 ninv = 0u - qcf_inv32(n[0]);
    du128 nv = 0;
#pragma unroll
    for (int i = LN - 1; i >= 0; --i) {
        nv = (nv << 32) | n[i];
    }
    u64 tested = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_claim = (*(volatile u32*)&prm.hits->stop) ? ~0ull : atomicAdd(prm.work, 1ull);
        }
        __syncthreads();
        const u64 srow = s_claim;
        if (srow >= n_srows) {
            break;
        }
        for (u32 e = threadIdx.x; e < QCF_GCD_ROWS * n_a; e += blockDim.x) {
            const u32 rr = e / n_a;
            const u32 i = e - rr * n_a;
            const u64 row = srow * QCF_GCD_ROWS + rr;
            if (row >= prm.n_rows) {
                break;
            }
            u64 c = row;
            u32 bj = 0;
            if (n_b > 1u) {
                c = __umul64hi(row, prm.nb_magic);
                bj = s_b[(u32)(row - c * n_b)];
            }
            u32 g0s = s_a[i] + bj;
            g0s = (g0s >= P) ? (g0s - P) : g0s;
            u32 g[LG];
            {
                const u64 lo = (u64)(u32)c * P;
                const u64 hi = (u64)(u32)(c >> 32) * P + (lo >> 32);
                const u32 prod[3] = { (u32)lo, (u32)hi, (u32)(hi >> 32) };
                u64 cy = g0s;
#pragma unroll
                for (int k = 0; k < LG; ++k) {
                    cy += (u64)prm.base[k] + prod[k];
                    g[k] = (u32)cy;
                    cy >>= 32;
                }
            }
            for (u32 k = 0; k < kc;) {
                const u32 run = (kc - k < QCF_GCD_BATCH) ? (kc - k) : QCF_GCD_BATCH;
                u32 g_start[LG], acc[LN], acc2[LN];
#pragma unroll
                for (int q = 0; q < LG; ++q) {
                    g_start[q] = g[q];
                }
#pragma unroll
                for (int q = 0; q < LN; ++q) {
                    acc[q] = (q == 0) ? 1u : 0u;
                    acc2[q] = (q == 0) ? 1u : 0u;
                }
                // two accumulators take alternate guesses: a Montgomery product is one long dependent chain (rows of multiplies
                // joined by carry chains), and two independent ones per thread hide each other's latency
                u32 r = 0;
                for (; r + 2u <= run; r += 2u) {
                    qcf_montmul_acc<LN, LG, LAZY>(acc, g, n, ninv);
                    qcf_addn<LG>(g, stride);
                    qcf_montmul_acc<LN, LG, LAZY>(acc2, g, n, ninv);
                    qcf_addn<LG>(g, stride);
                }
                if (r < run) {
                    qcf_montmul_acc<LN, LG, LAZY>(acc, g, n, ninv);
                    qcf_addn<LG>(g, stride);
                }
                du128 av = 0, av2 = 0;
#pragma unroll
                for (int q = LN - 1; q >= 0; --q) {
                    av = (av << 32) | acc[q];
                    av2 = (av2 << 32) | acc2[q];
                }
                if (qcf_gcd_not_one_128(nv, av) || qcf_gcd_not_one_128(nv, av2)) { // rare: some guess of this block shares a factor with N
                    u32 h[LG];
#pragma unroll
                    for (int q = 0; q < LG; ++q) {
                        h[q] = g_start[q];
                    }
                    for (u32 rr = 0; rr < run; ++rr) {
                        if (qcf_common_factor<LN, LG>(n, h) && !qcf_hits_extra<LG>(prm, h)) {
                            qcf_gcd_report<LG>(prm.hits, h);
                        }
                        qcf_addn<LG>(h, stride);
                    }
                }
                k += run;
            }
            tested += (prm.count && prm.n_extra) ? qcf_chain_count_extra(prm, c, g0s) : kc;
        }
    }
    if (prm.count) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            tested += __shfl_down_sync(0xffffffffu, tested, off);
        }
        if ((threadIdx.x & 31u) == 0u) {
            atomicAdd((unsigned long long*)&prm.hits->tested, (unsigned long long)tested);
        }
    }
}

template <int LN, int LG, bool LAZY> static cudaError_t qcf_launch_gcd_chain_t(const FilterParams& prm, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
    static QcfLaunchCache cache;
    int per_sm = 0;
    const cudaError_t e = qcf_resident_blocks(qcf_gcd_chain_kernel<LN, LG, LAZY>, cache, QCF_GCD_BLOCK, smem, &per_sm);
    if (e != cudaSuccess) {
        return e;
    }
    const u64 srows = (prm.n_rows + QCF_GCD_ROWS - 1u) / QCF_GCD_ROWS;
    const u64 blocks = (u64)sm_count * (u64)(per_sm > 0 ? per_sm : 1);
    qcf_gcd_chain_kernel<LN, LG, LAZY><<<(unsigned)(blocks < srows ? blocks : srows), QCF_GCD_BLOCK, smem, stream>>>(prm);
    return cudaGetLastError();
}

template <int LN, int LG> static cudaError_t qcf_launch_gcd_rows_t(const ExactParams& prm, bool count, int grid, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
    static QcfLaunchCache cache;
    int per_sm = 0;
    const cudaError_t e = qcf_resident_blocks(qcf_gcd_rows_kernel<LN, LG>, cache, QCF_BLOCK, smem, &per_sm);
    if (e != cudaSuccess) {
        return e;
    }
    qcf_gcd_rows_kernel<LN, LG><<<grid, QCF_BLOCK, smem, stream>>>(prm, count ? 1u : 0u);
    return cudaGetLastError();
}

#define QCF_GCD_CASES                                                                                                  \
    QCF_CASE(1, 1) QCF_CASE(2, 1) QCF_CASE(2, 2) QCF_CASE(3, 1) QCF_CASE(3, 2) QCF_CASE(3, 3) QCF_CASE(4, 1)            \
    QCF_CASE(4, 2) QCF_CASE(4, 3)

cudaError_t qcf_launch_gcd_chain(const FilterParams& prm, int ln, int lg, bool lazy, int sm_count, cudaStream_t stream)
{
#define QCF_CASE(LN, LG)                                                                                               \
    if ((ln == LN) && (lg == LG)) {                                                                                    \
        return lazy ? qcf_launch_gcd_chain_t<LN, LG, true>(prm, sm_count, stream)                                      \
                    : qcf_launch_gcd_chain_t<LN, LG, false>(prm, sm_count, stream);                                    \
    }
    QCF_GCD_CASES
#undef QCF_CASE
    return QCF_NO_KERNEL;
}

cudaError_t qcf_launch_gcd_rows(const ExactParams& prm, int ln, int lg, bool count, int grid, cudaStream_t stream)
{
#define QCF_CASE(LN, LG)                                                                                               \
    if ((ln == LN) && (lg == LG)) {                                                                                    \
        return qcf_launch_gcd_rows_t<LN, LG>(prm, count, grid, stream);                                                \
    }
    QCF_GCD_CASES
#undef QCF_CASE
    return QCF_NO_KERNEL;
}
