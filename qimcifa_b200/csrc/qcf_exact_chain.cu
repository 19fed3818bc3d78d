// Exact kernel, chain form: one exact Hensel/Montgomery divisibility test per guess (reference include/qimcifa.hpp:369),
// with the guesses of a thread taken along an arithmetic progression g_k = g_0 + k*S, S = period << tprime, exactly as
// the filter kernel lays them out (qcf_filter.cu) -- but every guess takes the exact test, there is no filter.
//
// What the progression buys: the index->guess map is one multi-limb add per guess, and when 2^16 | S the 2-adic
// reciprocal g^-1 mod 2^32 is carried from guess to guess by ONE Newton step x <- x*(2 - g*x)  (g_{k+1}*x_k = 1 + S*x_k
// has 16 trailing zero bits after the 1, a Newton step doubles that to 32): the "precomputed reciprocal" of the design.
// Per guess: QN*(1 + LG) multiply instructions for the reduction + 2 for the reciprocal.
#include <cstdio>
#include <cstdlib>

#include "qcf_device.cuh"
#include "qcf_internal.h"

#define QCF_XCHAIN_BLOCK 128
#define QCF_XCHAIN_ROWS 4

// by value: a reference to the guess array would force the guesses of every trip into local memory.  Multiples of the
// level's extra prime (29 at level 10) are reported like any divisor and dropped by the host (run_indices): a check here,
// even out of line, costs the chain loop registers at its 32-register cap.
__device__ __noinline__ void qcf_xchain_report_v(QcfHits* hits, u32 g0, u32 g1, u32 g2)
{
    const u32 slot = atomicAdd(&hits->count, 1u);
    if (slot < QCF_MAX_HITS) {
        hits->g[slot][0] = g0;
        hits->g[slot][1] = g1;
        hits->g[slot][2] = g2;
        hits->g[slot][3] = 0u;
    }
}

template <int LG> __device__ __forceinline__ void qcf_xchain_report(QcfHits* hits, const u32 (&g)[LG])
{
    qcf_xchain_report_v(hits, g[0], (LG > 1) ? g[(LG > 1) ? 1 : 0] : 0u, (LG > 2) ? g[(LG > 2) ? 2 : 0] : 0u);
}

template <int LN, int LG, int QN, bool INCR> __global__ void __launch_bounds__(QCF_XCHAIN_BLOCK, 16) qcf_exact_chain_kernel(const __grid_constant__ FilterParams prm)
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
    const u64 n_srows = (prm.n_rows + QCF_XCHAIN_ROWS - 1u) / QCF_XCHAIN_ROWS;
    u32 n[LN];
#pragma unroll
    for (int i = 0; i < LN; ++i) {
        n[i] = prm.n[i];
    }
    u32 stride[LG];
#pragma unroll
    for (int i = 0; i < LG; ++i) {
        stride[i] = prm.stride[i];
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
        for (u32 e = threadIdx.x; e < QCF_XCHAIN_ROWS * n_a; e += blockDim.x) {
            const u32 rr = e / n_a;
            const u32 i = e - rr * n_a;
            const u64 row = srow * QCF_XCHAIN_ROWS + rr;
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
            // g = base + c*P + g0s
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
            u32 x = qcf_inv32(g[0]);
            // V guesses per trip: their reductions are independent dependency chains, the hit branch is taken once per trip
            constexpr u32 V = 4;
            u32 k = 0;
            for (; k + V <= kc; k += V) {
                u32 gs[V][LG], xs[V];
#pragma unroll
                for (u32 v = 0; v < V; ++v) {
#pragma unroll
                    for (int q = 0; q < LG; ++q) {
                        gs[v][q] = g[q];
                    }
                    xs[v] = x;
                    qcf_addn<LG>(g, stride);
                    x = INCR ? (x * (2u - g[0] * x)) : qcf_inv32(g[0]);
                }
                bool hit[V];
                bool any = false;
#pragma unroll
                for (u32 v = 0; v < V; ++v) {
                    if ((LG == 2) && (QN == 2) && (LN >= 3)) {
                        hit[v] = qcf_divides_q2g2<LN>(n, gs[v][0], gs[v][LG - 1], 0u - xs[v]);
                    } else if (LG == 2) {
                        hit[v] = qcf_divides_ninv2<LN, QN>(n, gs[v][0], gs[v][LG - 1], 0u - xs[v]);
                    } else {
                        hit[v] = qcf_divides_ninv<LN, LG, QN>(n, gs[v], 0u - xs[v]);
                    }
                    any = any || hit[v];
                }
                if (any) {
#pragma unroll
                    for (u32 v = 0; v < V; ++v) {
                        if (hit[v]) {
                            qcf_xchain_report<LG>(prm.hits, gs[v]);
                        }
                    }
                }
            }
            for (; k < kc; ++k) {
                bool hit;
                if (LG == 2) {
                    hit = qcf_divides_ninv2<LN, QN>(n, g[0], g[LG - 1], 0u - x);
                } else {
                    hit = qcf_divides_ninv<LN, LG, QN>(n, g, 0u - x);
                }
                if (hit) {
                    qcf_xchain_report<LG>(prm.hits, g);
                }
                qcf_addn<LG>(g, stride);
                x = INCR ? (x * (2u - g[0] * x)) : qcf_inv32(g[0]);
            }
            tested += kc; // at level 10 this counts the level-9 guesses of the chain: the host subtracts the multiples of 29
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

template <int LN, int LG, int QN> static cudaError_t qcf_launch_xchain_t(const FilterParams& prm, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
    const bool incr = prm.tprime >= 15u; // 2^16 | S: one Newton step carries the reciprocal
    int per_sm = 0;
    cudaError_t e;
    if (incr) {
        cudaFuncSetAttribute(qcf_exact_chain_kernel<LN, LG, QN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qcf_exact_chain_kernel<LN, LG, QN, true>, QCF_XCHAIN_BLOCK, smem);
    } else {
        cudaFuncSetAttribute(qcf_exact_chain_kernel<LN, LG, QN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qcf_exact_chain_kernel<LN, LG, QN, false>, QCF_XCHAIN_BLOCK, smem);
    }
    if (e != cudaSuccess) {
        return e;
    }
    const u64 srows = (prm.n_rows + QCF_XCHAIN_ROWS - 1u) / QCF_XCHAIN_ROWS;
    const u64 blocks = (u64)sm_count * (u64)(per_sm > 0 ? per_sm : 1);
    const unsigned grid = (unsigned)(blocks < srows ? blocks : srows);
    if (incr) {
        qcf_exact_chain_kernel<LN, LG, QN, true><<<grid, QCF_XCHAIN_BLOCK, smem, stream>>>(prm);
    } else {
        qcf_exact_chain_kernel<LN, LG, QN, false><<<grid, QCF_XCHAIN_BLOCK, smem, stream>>>(prm);
    }
    return cudaGetLastError();
}

cudaError_t qcf_launch_exact_chain(const FilterParams& prm, int ln, int lg, int qn, int sm_count, cudaStream_t stream)
{
#define QCF_CASE(LN, LG, QN)                                                                                           \
    if ((ln == LN) && (lg == LG) && (qn == QN)) {                                                                      \
        return qcf_launch_xchain_t<LN, LG, QN>(prm, sm_count, stream);                                                 \
    }
    QCF_CASE(2, 1, 1) QCF_CASE(2, 1, 2) QCF_CASE(2, 2, 1) QCF_CASE(2, 2, 2)
    QCF_CASE(3, 1, 2) QCF_CASE(3, 1, 3) QCF_CASE(3, 2, 1) QCF_CASE(3, 2, 2) QCF_CASE(3, 2, 3)
    QCF_CASE(4, 1, 3) QCF_CASE(4, 1, 4) QCF_CASE(4, 2, 2) QCF_CASE(4, 2, 3) QCF_CASE(4, 2, 4) QCF_CASE(4, 3, 1) QCF_CASE(4, 3, 2)
#undef QCF_CASE
    return cudaErrorInvalidValue;
}
