// Internal declarations shared by the host side (qcf_api.cu) and the kernels (qcf_*.cu).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <mutex>

#include "qimcifa_cuda.h"

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

#define QCF_BLOCK 480 // threads per sweep block: one wheel-11 residue column each (480 = phi(2310))

// Launches of one sweep call are dealt round-robin to this many CUDA streams ("lanes"), so that the small launches of a
// call (leftovers, ragged ends, the narrow octaves far below sqrt(N)) overlap each other and the tail of the big one.
#define QCF_LANES 4

#define QCF_WORK_RING 64
static_assert(QCF_WORK_RING % QCF_LANES == 0, "a ring entry must always be reused on the stream of its previous user");

// Device-side hit record (one per context).  The host reads back the first QCF_HITS_D2H_BYTES after every call; the ring
// of row-group counters and the dispenser word behind them never leave the device.
struct QcfHits {
    u32 count;
    u32 stop;    // found flag polled by running kernels (finish(), reference include/qimcifa.hpp:102-105)
    u32 g[QCF_MAX_HITS][4];
    u64 tested;  // device-side tested-guess counter (QCF_COUNT_ON_DEVICE)
    u32 overflow; // survivors the filter launches of this call tested on the spot because the queue was full (a statistic)
    u32 pad_;
    // next unclaimed row group of each chain launch of the current call (dynamic distribution): a ring, zeroed ONCE per call
    // with the counters above; launch number i takes entry i % QCF_WORK_RING (and re-zeroes it on its own stream when the
    // ring wraps -- the previous user of the entry ran on the same stream, QCF_WORK_RING being a multiple of QCF_LANES)
    unsigned long long work[QCF_WORK_RING];
    unsigned long long dispenser; // batches handed out so far from this context's shared dispenser (qcf_peer_dispense)
};
#define QCF_HITS_D2H_BYTES (offsetof(QcfHits, work))

// Slot of a new hit in the record, or >= QCF_MAX_HITS when it is full.  The counter is the exact number of hits (n_hits of
// the result) up to 2^31 and saturates there instead of wrapping (common-factor mode on an N with a small prime factor
// reports a guess in every few), so "count > QCF_MAX_HITS" always means "the record is incomplete".
__device__ __forceinline__ u32 qcf_hit_slot(QcfHits* hits)
{
    if (*(volatile u32*)&hits->count >= 0x80000000u) {
        return ~0u;
    }
    return atomicAdd(&hits->count, 1u);
}

// Per-level wheel tables, resident in HBM (tiny) and staged into shared memory by every block.
//   guess = m * period + ((A[i] + B[j]) mod period),  i < n_a, j < n_b
struct QcfTables {
    u32 period; // product of the level's primes
    u32 n_a;    // 480 (L5) or 5760 (L>=6)
    u32 n_b;    // 1 (L5, L6), 16, 288, 6336 (L7..L9)
    u32* d_a;   // device
    u32* d_b;   // device
};

// One launch of the exact kernel covers values [v_lo, v_hi] (inclusive):
// periods m_lo..m_lo+m_span, rows = (m_span+1)*n_b, each row = n_a guesses.
struct ExactParams {
    u32 n[4];        // N, little-endian limbs
    u32 base[3];     // m_lo * period, little-endian limbs
    u32 v_lo[3];     // inclusive value bounds, checked only in the first/last period
    u32 v_hi[3];
    u32 period, n_a, n_b;
    u32 extra_prime; // 29 at level 10 (the tables are the level-9 ones): its multiples are neither counted nor reported
    u64 m_span;      // last period offset (number of periods - 1)
    u64 n_rows;      // (m_span + 1) * n_b
    const u32* d_a;
    const u32* d_b;
    QcfHits* hits;
};

// Test / calibration switches (qcf_set_option, include/qimcifa_cuda.h): one process-wide record, read once at the entry
// of an ABI call.  Nothing here is needed in normal use, and nothing is read from the environment.
struct QcfOptions {
    int xchain_two_stage = 1; // exact chain kernel: low limb first, full test out of line (1) or the full reduction per guess (0)
    int chain_min_log2 = 19;  // smallest run of periods (log2) that takes the chain form of the exact / common-factor kernels
    int force_d = 0, force_tp = 0, force_j = 0, force_pad = 0; // planner search restricted to this shape (0 = free)
    int force_drift = -1;     // -1 free, 0 windows without the linear cofactor drift, 1 with it
    int force_packed = 1;     // 0: degree-2 launches never take the packed 16-bit form
    unsigned epoch_trips = 0; // trips between two drains of the survivor queue (0 = sized from the survivor rate)
    int debug = 0;            // plans and launches on stderr
    int debug_stats = 0;      // survivor statistics per filter launch (serialises launches)
    double cost[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // planner cost constants (0 = the fitted value)
};
QcfOptions qcf_options_snapshot();

// Launchers (defined next to the kernels).
cudaError_t qcf_launch_exact(const ExactParams& prm, int ln, int lg, int qn, bool count, int grid, cudaStream_t stream);
cudaError_t qcf_launch_build_tables(int level, QcfTables* t, cudaStream_t stream);
cudaError_t qcf_launch_divisible(const u32* n, int ln, const u32* d_g, int lg, u64 count, uint8_t* d_out, cudaStream_t stream);
cudaError_t qcf_launch_microbench(int kind, u64 iters, int grid, int block, u32* d_sink, cudaStream_t stream);

// One launch of the filter kernel covers M whole wheel periods starting at value `base`:
//   guess = base + (c + (k << tprime)) * period + ((A[i] + B[j]) mod period),  c < 2^tprime,  k < K
// Every (c, i, j) is one arithmetic progression ("chain") with stride S = period << tprime; along a chain the
// aligned 2-adic quotient  Z_k = ((N * g_k^-1 - c_lo) << align) mod 2^64  is a polynomial of degree `degree`
// in k, stepped by finite differences on its high 32 bits only.  A guess survives when Z_hi <= bound.
#define QCF_MAX_EXTRA 5 // level-5 chains at level 10: 13, 17, 19, 23, 29
#define QCF_FILTER_MAX_SEG 32
struct FilterParams {
    u32 n[4];      // N, little-endian limbs (exact test of survivors)
    u32 base[3];   // value of the first period start (m_lo * period)
    u32 period, n_a, n_b;
    u32 tprime;    // log2 of the chain stride in periods
    u32 align;     // left shift applied to the quotient (64 - w)
    u32 bound;     // pass threshold of segment 0 (debug print only; the kernel uses seg_bound[])
    u32 slack;     // (steps walked) - 1, added to the initial high limb
    u32 steps;     // K: guesses per chain that belong to the launch; periods = K << tprime.  A chain WALKS whole trips,
                   // sum(seg_trips) * unroll >= K steps; the steps past K are never verified, reported or counted
    // A chain runs monotonically through the launch's value range, so it is cut into n_seg segments of seg_trips[j]
    // trips; segment j has its own cofactor window [c_lo_j, c_hi_j] (about n_seg times narrower than the launch's):
    // entering it, the tracked limb is re-based by seg_shift[j] and compared against seg_bound[j].
    u32 n_seg;
    u32 rows_per_group; // rows a block claims at a time: 4 (every thread gets whole chains), 1 when rows are scarce
    u32 epoch_trips; // trips between two drains of the survivor queue (sized by the host from the survivor rate)
    u32 seg_trips[QCF_FILTER_MAX_SEG];
    u32 seg_shift[QCF_FILTER_MAX_SEG];
    u32 seg_bound[QCF_FILTER_MAX_SEG];
    // Linear cofactor drift: inside segment j the quotient is compared against a LINE  l_j(k) = beta_j - lambda_j * (k - ka_j)
    // that follows the cofactor N/g_k down the chain, not against a constant floor, so a segment's window is the cofactor's
    // curvature over the segment instead of its whole movement.  mu0 = lambda_0 << align is added to the chain's first
    // difference at setup; entering segment j >= 1 the 64-bit first difference moves by seg_nu[j] = (lambda_j - lambda_{j-1}) << align.
    u32 drift;     // 0: constant floors (every seg_nu is 0 and the kernel skips the slope update)
    u32 seg_nu_lo[QCF_FILTER_MAX_SEG];
    u32 seg_nu_hi[QCF_FILTER_MAX_SEG];
    u64 mu0;
    // per-chain offset correction (qcf_filter_rho32 / qcf_filter_rho_corr, qcf_filter_math.cuh); rho = 0: not used
    u32 rho;
    u32 pinv32;    // floor(2^32 / period)
    u64 sigma0;    // bound on one step's cofactor drift at the launch's first step
    u32 seg_kappa[QCF_FILTER_MAX_SEG];
    // chain-draw random mode (QCF_RANDOM_CHAINS): the launch walks rand_count row groups DRAWN by Philox (sample numbers
    // rand_first .. rand_first + rand_count - 1 of counter block rand_block) instead of all of them in order
    u32 rand_on;
    u32 rand_key0, rand_key1;
    u32 rand_block_lo, rand_block_hi;
    u32 rand_first, rand_count;
    u32 count;     // 1: count tested guesses on the device
    u32 stride[3]; // period << tprime, little-endian limbs (exact chain kernel)
    u32 n_extra;   // wheel primes of the level that the chain tables do not remove (superset mode; 29 at level 10)
    u32 extra[QCF_MAX_EXTRA];
    u64 n12p;      // exact chain kernel, two-stage loop: (floor(N / 2^32) + [N mod 2^32 != 0]) mod 2^64 (qcf_lowlimb_q2g2_q0)
    u64 n64;       // N mod 2^64
    u64 c_lo;      // the reference line's value at step 0, l_0(0), mod 2^64 (without drift: the cofactor floor of segment 0)
    u64 n_rows;    // n_b << tprime
    u64 nb_magic;  // floor(2^64 / n_b) + 1: row / n_b == umul64hi(row, nb_magic)
    const u32* d_a;
    const u32* d_b;
    QcfHits* hits;
    unsigned long long* work; // row-group counter of this launch (an entry of hits->work[])
    u32* dbg;      // QCF_DEBUG_STATS only: [0] = exact tests of survivors, [2 + (chain & 0xFFFF)] = the same per chain
};

#include "qcf_trip.h" // QCF_FILTER_UNROLL_FOR: guesses per trip of the filter kernel, one definition for host and device
cudaError_t qcf_launch_filter(const FilterParams& prm, int degree, bool packed, int ln, int lg, int grid, cudaStream_t stream);

// Random-guess mode: counters [counter_lo, counter_lo + budget), Philox4x32-10 keyed by the seed.
struct RandomParams {
    u32 n[4];
    u32 period, n_a, n_b;
    u32 key0, key1;
    u64 periods;     // M: whole periods below sqrt(N)
    u64 counter_lo, budget;
    const u32* d_a;
    const u32* d_b;
    QcfHits* hits;
    u32* dump;       // optional: guesses out (4 limbs each), no divisibility test
    u32 count;
};
cudaError_t qcf_launch_random(const RandomParams& prm, int ln, int lg, int grid, cudaStream_t stream);

// Exact kernel in chain form (qcf_exact_chain.cu): same layout of guesses as the filter kernel (FilterParams: base, tprime,
// steps, tables), one exact divisibility test per guess.
cudaError_t qcf_launch_exact_chain(const FilterParams& prm, int ln, int lg, int qn, int sm_count, bool two_stage, cudaStream_t stream);
// "no instantiation for this shape": the caller tries the next wider one.  A value no CUDA call returns for a launch.
#define QCF_NO_KERNEL cudaErrorNotSupported

// Common-factor mode (qcf_gcd.cu): gcd(guess, N) != 1, N odd.  Row form for ragged ends and short intervals, chain form
// (Montgomery product of a block of guesses, one gcd per block) for whole periods.
cudaError_t qcf_launch_gcd_rows(const ExactParams& prm, int ln, int lg, bool count, int grid, cudaStream_t stream);
cudaError_t qcf_launch_gcd_chain(const FilterParams& prm, int ln, int lg, bool lazy, int sm_count, cudaStream_t stream);

// Writes 1 into the found-flag word of every connected peer GPU (system-scope atomics over NVLink peer memory).
cudaError_t qcf_launch_signal_peers(u32* const* d_peer_flags, int n_peers, cudaStream_t stream);
// *out = atomicAdd_system(word, take): the dispenser word may live in a peer GPU's memory (NVLink).
cudaError_t qcf_launch_dispense(unsigned long long* word, unsigned long long take, unsigned long long* out, cudaStream_t stream);

// ---- launch bookkeeping shared by the chain launchers -------------------------------------------------------------------
// The dynamic-shared-memory attribute of a kernel is per (function, device) and shared by all threads and contexts, so it
// is raised ONCE per device to the largest table set any level needs (level 9: 5760 + 6336 words) and never lowered;
// resident blocks per SM are looked up once per (device, shared-memory size).  One cache object per kernel instantiation
// (a function-local static of the launcher), guarded by its own mutex: two contexts on one GPU driven from different
// threads at different levels cannot leave each other a stale attribute.
#define QCF_MAX_TABLE_SMEM ((size_t)4 * (5760 + 6336))
struct QcfLaunchCache {
    std::mutex mu;
    bool attr_done[64] = {};
    struct Entry {
        int dev;
        size_t smem;
        int per_sm;
    } e[16];
    int n = 0;
};
template <typename K> cudaError_t qcf_resident_blocks(K kernel, QcfLaunchCache& c, int block, size_t smem, int* per_sm)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) {
        return err;
    }
    std::lock_guard<std::mutex> lock(c.mu);
    if ((dev < 64) && !c.attr_done[dev]) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCF_MAX_TABLE_SMEM);
        if (err != cudaSuccess) {
            return err;
        }
        c.attr_done[dev] = true;
    }
    for (int i = 0; i < c.n; ++i) {
        if ((c.e[i].dev == dev) && (c.e[i].smem == smem)) {
            *per_sm = c.e[i].per_sm;
            return cudaSuccess;
        }
    }
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kernel, block, smem);
    if (err != cudaSuccess) {
        return err;
    }
    if (c.n < 16) {
        c.e[c.n++] = { dev, smem, *per_sm };
    }
    return cudaSuccess;
}
