// Internal declarations shared by the host side (qcf_api.cu) and the kernels (qcf_*.cu).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "qimcifa_cuda.h"

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

#define QCF_BLOCK 480 // threads per sweep block: one wheel-11 residue column each (480 = phi(2310))

// Launches of one sweep call are dealt round-robin to this many CUDA streams ("lanes"), so that the small launches of a
// call (leftovers, ragged ends, the narrow octaves far below sqrt(N)) overlap each other and the tail of the big one.
#define QCF_LANES 4

// Device-side hit record (one per context).
struct QcfHits {
    u32 count;
    u32 stop;    // found flag polled by running kernels (finish(), reference include/qimcifa.hpp:102-105)
    u64 tested;  // device-side tested-guess counter (QCF_COUNT_ON_DEVICE)
    u32 overflow; // a filter launch dropped survivors (queue full): the host repeats its range with the exact kernel
    u32 pad_;
    unsigned long long work[QCF_LANES]; // next unclaimed row group of the launch running on each lane (dynamic distribution)
    u32 g[QCF_MAX_HITS][4];
    unsigned long long dispenser; // batches handed out so far from this context's shared dispenser (qcf_peer_dispense)
};

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
    u32 count;     // 1: count tested guesses on the device
    u32 stride[3]; // period << tprime, little-endian limbs (exact chain kernel)
    u32 n_extra;   // wheel primes of the level that the chain tables do not remove (superset mode; 29 at level 10)
    u32 extra[QCF_MAX_EXTRA];
    u64 n64;       // N mod 2^64
    u64 c_lo;      // floor(N / (end of range)) mod 2^64
    u64 n_rows;    // n_b << tprime
    u64 nb_magic;  // floor(2^64 / n_b) + 1: row / n_b == umul64hi(row, nb_magic)
    const u32* d_a;
    const u32* d_b;
    QcfHits* hits;
    unsigned long long* work; // row-group counter of this launch (&hits->work[lane])
    u32* dbg;      // QCF_DEBUG_STATS only: [0] = exact tests of survivors, [2 + (chain & 0xFFFF)] = the same per chain
};

// guesses per trip of the filter kernel: 32 at degree <= 2, 16 at degree 3 (the same definition as in qcf_filter_math.cuh,
// which the host build of the chain arithmetic includes without this header)
#ifndef QCF_FILTER_UNROLL_FOR
#define QCF_FILTER_UNROLL_FOR(D) (((D) >= 3) ? 16 : 32)
#endif
cudaError_t qcf_launch_filter(const FilterParams& prm, int degree, int ln, int lg, int grid, cudaStream_t stream);

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
cudaError_t qcf_launch_exact_chain(const FilterParams& prm, int ln, int lg, int qn, int sm_count, cudaStream_t stream);

// Common-factor mode (qcf_gcd.cu): gcd(guess, N) != 1, N odd.  Row form for ragged ends and short intervals, chain form
// (Montgomery product of a block of guesses, one gcd per block) for whole periods.
cudaError_t qcf_launch_gcd_rows(const ExactParams& prm, int ln, int lg, bool count, int grid, cudaStream_t stream);
cudaError_t qcf_launch_gcd_chain(const FilterParams& prm, int ln, int lg, int sm_count, cudaStream_t stream);

// Writes 1 into the found-flag word of every connected peer GPU (system-scope atomics over NVLink peer memory).
cudaError_t qcf_launch_signal_peers(u32* const* d_peer_flags, int n_peers, cudaStream_t stream);
// *out = atomicAdd_system(word, take): the dispenser word may live in a peer GPU's memory (NVLink).
cudaError_t qcf_launch_dispense(unsigned long long* word, unsigned long long take, unsigned long long* out, cudaStream_t stream);
