// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; constants of the paper):
// the counter-based generator of the random-guess Monte-Carlo mode (reference README.md:25 describes the mode; the snapshot
// has no code for it).  Host and device: the host draws the window of a counter block with it, the kernels draw guesses
// and row groups.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define QCF_HD __host__ __device__ __forceinline__
#else
#define QCF_HD inline
#endif

QCF_HD void qcf_philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// floor(x * range / 2^64) for the 64-bit draw x = r1:r0: uniform in [0, range)
QCF_HD uint64_t qcf_philox_below(uint32_t r0, uint32_t r1, uint64_t range)
{
    const uint64_t x = ((uint64_t)r1 << 32) | r0;
#if defined(__CUDA_ARCH__)
    return __umul64hi(x, range);
#else
    return (uint64_t)(((unsigned __int128)x * range) >> 64);
#endif
}

// Chain-draw mode (QCF_RANDOM_CHAINS, include/qimcifa_cuda.h): the window of counter block b and the row group of its
// j-th sample.  The third and fourth counter words keep the two streams apart.
#define QCF_RANDOM_BLOCK_LOG2 34
#define QCF_RANDOM_WINDOW_BATCHES 4096u
QCF_HD uint64_t qcf_random_window_of(uint64_t seed, uint64_t block, uint64_t n_windows)
{
    uint32_t c[4] = { (uint32_t)block, (uint32_t)(block >> 32), 0x51434657u, 0u };
    qcf_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return qcf_philox_below(c[0], c[1], n_windows);
}
QCF_HD uint64_t qcf_random_row_group_of(uint32_t key0, uint32_t key1, uint32_t block_lo, uint32_t block_hi, uint32_t sample, uint64_t n_row_groups)
{
    uint32_t c[4] = { block_lo, block_hi, sample, 1u };
    qcf_philox4x32_10(c, key0, key1);
    return qcf_philox_below(c[0], c[1], n_row_groups);
}
