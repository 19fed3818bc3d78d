// Random-guess Monte-Carlo mode (reference README.md:25 describes it; the snapshot has no code for it:
// IS_RANDOM is a dead macro, include/common/config.h.in:6).  Counter-based Philox4x32-10 mapped through the
// same wheel tables as the sweep, one exact divisibility test per guess.
#include "qcf_device.cuh"
#include "qcf_internal.h"
#include "qcf_philox.cuh"

template <int LN, int LG> __global__ void __launch_bounds__(256) qcf_random_kernel(const RandomParams prm)
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
    u32 n[LN];
#pragma unroll
    for (int k = 0; k < LN; ++k) {
        n[k] = prm.n[k];
    }
    const u32 P = prm.period;
    u64 tested = 0;
    u32 iter = 0;
    for (u64 idx = blockIdx.x * (u64)blockDim.x + threadIdx.x; idx < prm.budget; idx += (u64)gridDim.x * blockDim.x) {
        const u64 ctr = prm.counter_lo + idx;
        u32 c[4] = { (u32)ctr, (u32)(ctr >> 32), 0u, 0u };
        qcf_philox4x32_10(c, prm.key0, prm.key1);
        const u64 m = __umul64hi(((u64)c[1] << 32) | c[0], prm.periods);
        const u32 ia = __umulhi(c[2], prm.n_a), ib = __umulhi(c[3], prm.n_b);
        u32 g0 = s_a[ia] + s_b[ib];
        g0 = (g0 >= P) ? (g0 - P) : g0;
        // g = m * P + g0  (up to 96 bits)
        const u64 lo = (u64)(u32)m * P;
        const u64 hi = (u64)(u32)(m >> 32) * P + (lo >> 32);
        const u32 prod[3] = { (u32)lo, (u32)hi, (u32)(hi >> 32) };
        u32 g[LG];
        u64 cy = g0;
#pragma unroll
        for (int k = 0; k < LG; ++k) {
            cy += prod[k];
            g[k] = (u32)cy;
            cy >>= 32;
        }
        if (prm.dump) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                prm.dump[idx * 4 + k] = (k < LG) ? g[k] : 0u;
            }
            continue;
        }
        bool is_one = (g[0] == 1u);
#pragma unroll
        for (int k = 1; k < LG; ++k) {
            is_one = is_one && (g[k] == 0u);
        }
        if (!is_one) {
            ++tested;
            if (qcf_divides<LN, LG, LN>(n, g)) {
                const u32 slot = qcf_hit_slot(prm.hits);
                if (slot < QCF_MAX_HITS) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        prm.hits->g[slot][k] = (k < LG) ? g[k] : 0u;
                    }
                }
            }
        }
        if (((++iter) & 255u) == 0u) {
            if (*(volatile u32*)&prm.hits->stop) {
                break;
            }
        }
    }
    if (!prm.dump) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            tested += __shfl_down_sync(0xffffffffu, tested, off);
        }
        if ((threadIdx.x & 31u) == 0u) {
            atomicAdd((unsigned long long*)&prm.hits->tested, (unsigned long long)tested);
        }
    }
}

cudaError_t qcf_launch_random(const RandomParams& prm, int ln, int lg, int grid, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
#define QCF_CASE(LN, LG)                                                                                               \
    if ((ln == LN) && (lg == LG)) {                                                                                    \
        static QcfLaunchCache cache;                                                                                   \
        int per_sm = 0;                                                                                                \
        const cudaError_t e = qcf_resident_blocks(qcf_random_kernel<LN, LG>, cache, 256, smem, &per_sm);               \
        if (e != cudaSuccess) {                                                                                        \
            return e;                                                                                                  \
        }                                                                                                              \
        qcf_random_kernel<LN, LG><<<grid, 256, smem, stream>>>(prm);                                                   \
        return cudaGetLastError();                                                                                     \
    }
    QCF_CASE(2, 1) QCF_CASE(2, 2) QCF_CASE(3, 1) QCF_CASE(3, 2) QCF_CASE(4, 1) QCF_CASE(4, 2) QCF_CASE(4, 3)
#undef QCF_CASE
    return QCF_NO_KERNEL;
}
