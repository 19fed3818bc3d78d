// Exact kernel, chain form: one exact Hensel/Montgomery divisibility test per guess (reference include/qimcifa.hpp:369),
// with the guesses of a thread taken along an arithmetic progression g_k = g_0 + k*S, S = period << tprime, exactly as
// the filter kernel lays them out (qcf_filter.cu) -- but every guess takes the exact test, there is no filter.
//
// What the progression buys: the index->guess map is one multi-limb add per guess, and when 2^16 | S the 2-adic
// reciprocal g^-1 mod 2^32 is LINEAR along the chain (qcf_ninv_delta, qcf_device.cuh: g_k^-1 = x0 - k*S*x0^2 exactly, the
// square of k*S*x0 vanishes mod 2^32), so it is carried from guess to guess by ONE ADD: the "precomputed reciprocal" of
// the design.  Per guess: QN*(1 + LG) multiply instructions for the reduction, none for the reciprocal.
//
// Two-stage form (TWO, the shape of every 65..128-bit sweep near sqrt(N): two-limb guess, two quotient words): the loop
// computes only limb 0 of the final t (qcf_lowlimb_q2g2, qcf_device.cuh: six multiply instructions, no carry chain above
// bit 64) and compares it with g0 -- necessary for g | N, passed by a non-divisor with probability 2^-32 -- and the
// trips in which some guess passes go out of line to the full test (qcf_xchain_second_stage).  Results are those of
// the full test; the negated reciprocal is carried directly (qcf_ninv_step), and the guesses of a trip are not kept:
// the second stage walks back from the current guess.
#include <cstdio>
#include <cstdlib>

#include "qcf_device.cuh"
#include "qcf_internal.h"

#define QCF_XCHAIN_BLOCK 128
#define QCF_XCHAIN_ROWS 4
// The two-stage loop is the default (QcfOptions::xchain_two_stage): 1.447e12 against 8.66e11 guesses/s on a B200
// (profiles/r01c_exact_chain_two_stage_ab.txt); qcf_set_option("xchain_two_stage", "0") selects the one-stage loop.

// by value: a reference to the guess array would force the guesses of every trip into local memory.  Multiples of the
// level's extra prime (29 at level 10) are reported like any divisor and dropped by the host (run_indices): a check here,
// even out of line, costs the chain loop registers at its 32-register cap.
// PATH keeps the two callers' copies apart: the second stage reaches the record through the parameter block's generic
// address, which would turn the one-stage kernels' global-space atomics and stores into generic ones too.
template <int PATH> __device__ __noinline__ void qcf_xchain_report_v(QcfHits* hits, u32 g0, u32 g1, u32 g2)
{
    const u32 slot = qcf_hit_slot(hits);
    if (slot < QCF_MAX_HITS) {
        hits->g[slot][0] = g0;
        hits->g[slot][1] = g1;
        hits->g[slot][2] = g2;
        hits->g[slot][3] = 0u;
    }
}

template <int LG> __device__ __forceinline__ void qcf_xchain_report(QcfHits* hits, const u32 (&g)[LG])
{
    qcf_xchain_report_v<0>(hits, g[0], (LG > 1) ? g[(LG > 1) ? 1 : 0] : 0u, (LG > 2) ? g[(LG > 2) ? 2 : 0] : 0u);
}

// Second stage of the two-stage form: `g` is the guess AFTER the `back` guesses of the trip in which some guess passed
// stage 1; each of them takes the full test (reciprocal from scratch) and true divisors are reported.  Out of line and
// fed from the parameter block, so that it costs the chain loop no registers: it runs once per ~2^30 guesses.
template <int LN> __device__ __noinline__ void qcf_xchain_second_stage(const FilterParams* prm, u32 g0, u32 g1, u32 back)
{
    u32 n[LN];
#pragma unroll
    for (int i = 0; i < LN; ++i) {
        n[i] = prm->n[i];
    }
    const u64 s = ((u64)prm->stride[1] << 32) | prm->stride[0];
    u64 h = (((u64)g1 << 32) | g0) - s * back; // two-limb guesses: the chain never leaves [0, 2^64)
    for (u32 v = 0; v < back; ++v) {
        const u32 h0 = (u32)h, h1 = (u32)(h >> 32);
        if (qcf_divides_q2g2<LN>(n, h0, h1, 0u - qcf_inv32(h0))) {
            qcf_xchain_report_v<1>(prm->hits, h0, h1, 0u);
        }
        h += s;
    }
}

// 16 blocks of 4 warps per SM at 32 registers (12 blocks at 40 registers measured slower: the loops are dependent chains of
// multiplies and occupancy hides them).
template <int LN, int LG, int QN, bool INCR, bool TWO> __global__ void __launch_bounds__(QCF_XCHAIN_BLOCK, TWO ? 12 : 16) qcf_exact_chain_kernel(const __grid_constant__ FilterParams prm)
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
            constexpr u32 V = 4;
            if constexpr (TWO) {
                static_assert(!TWO || ((LG == 2) && (QN == 2) && (LN >= 3)), "two-stage form: two-limb guess, two quotient words");
                // INCR (2^16 | stride): y = -g^-1 and q0 = n0 * y are linear along the chain (qcf_ninv_delta): one add each
                u32 y = 0u - qcf_inv32(g[0]);
                u32 q0 = n[0] * y;
                const u32 dy = INCR ? qcf_ninv_delta(y, stride[0]) : 0u;
                u32 dq = n[0] * dy;
#if defined(__CUDA_ARCH__)
                asm volatile("" : "+r"(dq)); // an opaque register: otherwise q0 + n0 * dy is re-formed as a multiply-add on the pipe that bounds the loop
#endif
                u32 left = kc; // one down-counter
                for (; left >= V; left -= V) {
                    bool any = false;
#pragma unroll
                    for (u32 v = 0; v < V; ++v) {
                        any |= qcf_lowlimb_q2g2_q0<LN>(n, q0, g[0], g[LG - 1], y); // not ||: no branch per guess
                        qcf_addn<LG>(g, stride);
                        if (INCR) {
                            y += dy;
                            q0 += dq;
                        } else {
                            y = 0u - qcf_inv32(g[0]);
                            q0 = n[0] * y;
                        }
                    }
                    if (any) {
                        qcf_xchain_second_stage<LN>(&prm, g[0], g[LG - 1], V);
                    }
                }
                for (; left; --left) {
                    const bool pass = qcf_lowlimb_q2g2_q0<LN>(n, q0, g[0], g[LG - 1], y);
                    qcf_addn<LG>(g, stride);
                    if (INCR) {
                        y += dy;
                        q0 += dq;
                    } else {
                        y = 0u - qcf_inv32(g[0]);
                        q0 = n[0] * y;
                    }
                    if (pass) {
                        qcf_xchain_second_stage<LN>(&prm, g[0], g[LG - 1], 1u);
                    }
                }
                tested += kc;
                continue;
            }
            u32 x = qcf_inv32(g[0]);
            const u32 dx = INCR ? stride[0] * x * x : 0u; // x_k = x_0 - k * S * x_0^2 (mod 2^32) when 2^16 | S: the reciprocal is linear along the chain
            // V guesses per trip: their reductions are independent dependency chains, the hit branch is taken once per trip
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
                    x = INCR ? (x - dx) : qcf_inv32(g[0]);
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
                x = INCR ? (x - dx) : qcf_inv32(g[0]);
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

template <int LN, int LG, int QN, bool INCR, bool TWO> static cudaError_t qcf_launch_xchain_v(const FilterParams& prm, int sm_count, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
    static QcfLaunchCache cache;
    int per_sm = 0;
    const cudaError_t e = qcf_resident_blocks(qcf_exact_chain_kernel<LN, LG, QN, INCR, TWO>, cache, QCF_XCHAIN_BLOCK, smem, &per_sm);
    if (e != cudaSuccess) {
        return e;
    }
    const u64 srows = (prm.n_rows + QCF_XCHAIN_ROWS - 1u) / QCF_XCHAIN_ROWS;
    const u64 blocks = (u64)sm_count * (u64)(per_sm > 0 ? per_sm : 1);
    const unsigned grid = (unsigned)(blocks < srows ? blocks : srows);
    qcf_exact_chain_kernel<LN, LG, QN, INCR, TWO><<<grid, QCF_XCHAIN_BLOCK, smem, stream>>>(prm);
    return cudaGetLastError();
}

template <int LN, int LG, int QN> static cudaError_t qcf_launch_xchain_t(const FilterParams& prm, int sm_count, bool two_stage, cudaStream_t stream)
{
    const bool incr = prm.tprime >= 15u; // 2^16 | S: one Newton step carries the reciprocal
    if constexpr ((LG == 2) && (QN == 2) && (LN >= 3)) {
        if (two_stage) { // qcf_set_option("xchain_two_stage"): A/B timing and the parity test of one loop against the other
            return incr ? qcf_launch_xchain_v<LN, LG, QN, true, true>(prm, sm_count, stream)
                        : qcf_launch_xchain_v<LN, LG, QN, false, true>(prm, sm_count, stream);
        }
    }
    return incr ? qcf_launch_xchain_v<LN, LG, QN, true, false>(prm, sm_count, stream)
                : qcf_launch_xchain_v<LN, LG, QN, false, false>(prm, sm_count, stream);
}

cudaError_t qcf_launch_exact_chain(const FilterParams& prm, int ln, int lg, int qn, int sm_count, bool two_stage, cudaStream_t stream)
{
#define QCF_CASE(LN, LG, QN)                                                                                           \
    if ((ln == LN) && (lg == LG) && (qn == QN)) {                                                                      \
        return qcf_launch_xchain_t<LN, LG, QN>(prm, sm_count, two_stage, stream);                                                \
    }
    QCF_CASE(2, 1, 1) QCF_CASE(2, 1, 2) QCF_CASE(2, 2, 1) QCF_CASE(2, 2, 2)
    QCF_CASE(3, 1, 2) QCF_CASE(3, 1, 3) QCF_CASE(3, 2, 1) QCF_CASE(3, 2, 2) QCF_CASE(3, 2, 3)
    QCF_CASE(4, 1, 3) QCF_CASE(4, 1, 4) QCF_CASE(4, 2, 2) QCF_CASE(4, 2, 3) QCF_CASE(4, 2, 4) QCF_CASE(4, 3, 1) QCF_CASE(4, 3, 2)
#undef QCF_CASE
    return QCF_NO_KERNEL;
}
