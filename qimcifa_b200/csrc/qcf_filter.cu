// Filter kernel: the fast path of the sweep (replaces the per-guess `toFactor % base`, reference
// include/qimcifa.hpp:369, by ~3 integer adds per guess plus an exact test of the rare survivors).
//
// Math (all exact, modulo 2^64).  For an odd guess g let  Q(g) = N * g^-1 mod 2^64.  If g | N then Q(g) is the
// cofactor C = N/g (mod 2^64), and C lies in [c_lo, c_hi] = [floor(N/v_end), floor(N/v_begin)] for every guess of
// the launch, so with a = align
//     Z(g) = ((Q(g) - c_lo) << a) mod 2^64  <=  (c_hi - c_lo) << a            (necessary for g | N).
// Along a chain g_k = g_0 + k*S with 2^t | S:  g_k^-1 = x * sum_i (-k*S*x)^i,  x = g_0^-1, and the terms with
// a + t*i >= 64 vanish, so Z(g_k) is a polynomial of degree D = ceil((64-a)/t) - 1 in k; its forward differences
// of order >= 2 are multiples of 2^32 (a + 2t >= 32), so only the HIGH limb of Z and of the differences changes
// from step to step, except for the carries out of the constant low-limb increment.  Those carries (at most one
// per step) are not tracked: the high limb is started K_max - 1 too high and the threshold widened by the same
// amount, which keeps the test a necessary condition.  A chain crosses the launch's value range monotonically, so it is
// cut into up to 32 segments of whole trips, each compared against its own (narrower) cofactor window: up to 5 more
// filter bits.  Survivors (probability ~ bound / 2^32 per guess) are queued in shared memory and take the exact
// Hensel/Montgomery test qcf_divides<> when the block drains the queue (or on the spot if the queue is full); a true
// divisor can never be rejected.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "qcf_chain_count.cuh"
#include "qcf_device.cuh"
#include "qcf_filter_math.cuh"
#include "qcf_internal.h"
#include "qcf_philox.cuh"

template <int LG> __device__ __forceinline__ void qcf_filter_report(QcfHits* hits, const u32 (&g)[LG])
{
    const u32 idx = qcf_hit_slot(hits);
    if (idx < QCF_MAX_HITS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hits->g[idx][k] = (k < LG) ? g[k] : 0u;
        }
    }
}

// Survivors are not tested where they are found: the walking lane appends (chain, step) to a queue in shared memory
// and walks on; the block drains the queue together at the end of an epoch (a run of trips sized by the host so that
// an epoch yields a few dozen survivors), one survivor per thread.  The chain loop therefore contains no call and no
// long divergent path -- a lane that finds a survivor delays its warp by a dozen instructions, not by an exact test.
#define QCF_FILTER_QCAP 256 // entries per queue; two queues alternate between epochs (see the barriers in the kernel)

// Exact test of one queued survivor: step k of chain (c, g0s); rebuilds the full-width guess.
template <int LN, int LG> __device__ __forceinline__ void qcf_filter_verify(const FilterParams& prm, u64 c, u32 k, u32 g0s)
{
    const u64 mp = c + ((u64)k << prm.tprime); // period offset inside the launch
    const u64 lo = (u64)(u32)mp * prm.period;
    const u64 hi = (u64)(u32)(mp >> 32) * prm.period + (lo >> 32);
    const u32 prod[3] = { (u32)lo, (u32)hi, (u32)(hi >> 32) };
    u32 g[LG];
    u64 cy = g0s;
#pragma unroll
    for (int i = 0; i < LG; ++i) {
        cy += (u64)prm.base[i] + prod[i];
        g[i] = (u32)cy;
        cy >>= 32;
    }
    // superset mode: the chains run over a lower wheel level; the level's remaining primes are applied here, so only
    // guesses of the requested level can ever be reported
    if (qcf_hits_extra<LG>(prm, g)) {
        return;
    }
    u32 n[LN];
#pragma unroll
    for (int i = 0; i < LN; ++i) {
        n[i] = prm.n[i];
    }
    if (qcf_divides<LN, LG, LN>(n, g)) {
        qcf_filter_report<LG>(prm.hits, g);
    }
}

// A survivor that finds the queue full is tested on the spot by the lane that met it.  Rare by construction (the epoch
// length is sized for ~48 of 256 entries), but not impossible: the planner's survivor rate is an expectation over
// pseudo-random quotients, and at degree 1 the tracked limb of a chain is an arithmetic progression whose increment
// takes few distinct values (a multiple of 2^(align + t - 32) plus a launch constant) -- a chain whose increment is tiny
// stays under the threshold for hundreds of consecutive steps.  The total number of survivors is unchanged (the start
// values are equidistributed), only their clustering is, so the cost of this path stays within the planner's estimate.
// QcfHits::overflow counts these survivors (a statistic: qcf_last_queue_spills).
template <int LN, int LG> __device__ __noinline__ void qcf_filter_spill(const FilterParams& prm, u32 c_lo, u32 c_hi, u32 k, u32 g0s)
{
    atomicAdd(&prm.hits->overflow, 1u);
    qcf_filter_verify<LN, LG>(prm, ((u64)c_hi << 32) | c_lo, k, g0s);
}

// Step `step` of the lane's chain passed the filter: append (chain, step) to the epoch's queue, or test it on the spot
// when the queue is full.  Steps past K pad the last trip: not part of the launch.
template <int LN, int LG> __device__ __forceinline__ void qcf_filter_queue(const FilterParams& prm, const uint4 ch, uint4* q, u32* qn, u32 step)
{
    if (step < ch.w) {
        const u32 slot = atomicAdd(qn, 1u);
        if (slot < QCF_FILTER_QCAP) {
            q[slot] = make_uint4(ch.x, ch.y, step, ch.z);
        } else {
            qcf_filter_spill<LN, LG>(prm, ch.x, ch.y, step, ch.z);
        }
    }
}

// A flagged trip holds a step at or below the bound in some lane: every lane walks the steps of ITS flagged classes (u mod 4,
// qcf_filter_math.cuh) -- a quarter of the trip -- and queues those at or below the bound; one round per class some lane
// still has (almost always one).  The whole warp walks together (a lane that walked alone never rejoined its warp); walking
// all U steps of the trip, as the first version did, made a survivor cost 8 700 lane-cycles.
// (zt, dt, d2) = the tracked limb and its differences at the trip's first step k0.
template <int LN, int LG, u32 U> __device__ __forceinline__ void qcf_filter_queue_classes(const FilterParams& prm, const uint4 ch, uint4* q, u32* qn, u32 zt, u32 dt,
    u32 d2, u32 cls, u32 bound, u32 k0)
{
    do {
        const u32 r = cls ? qcf_lowest_bit_index(cls) : 0u;
        u32 v, dv, ddv;
        qcf_filter_class_start(zt, dt, d2, r, v, dv, ddv);
#pragma unroll 1
        for (u32 u = r; u < U; u += 4u) {
            if (cls && (v <= bound)) {
                qcf_filter_queue<LN, LG>(prm, ch, q, qn, k0 + u);
            }
            v += dv;
            dv += ddv;
        }
        cls &= cls - 1u;
    } while (__any_sync(0xffffffffu, cls != 0u));
}

// End of an epoch, called by every thread of the block between two barriers: thread t tests entries t, t + 128, ...
template <int LN, int LG> __device__ __noinline__ void qcf_filter_drain(const FilterParams& prm, const uint4* q, u32 n)
{
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const uint4 e = q[i];
        qcf_filter_verify<LN, LG>(prm, ((u64)e.y << 32) | e.x, e.z, e.w);
    }
}

// End of an epoch: the block tests the queued survivors together.  Barrier 1: every append to queue `qsel` is done.
// Barrier 2: every thread has read the count and finished its share, so thread 0 may clear the counter; the next epoch
// appends to the other queue, whose counter was cleared one epoch ago, before thread 0 arrived at barrier 1.
template <int LN, int LG> __device__ __forceinline__ void qcf_filter_epoch_end(const FilterParams& prm, uint4 (*s_q)[QCF_FILTER_QCAP], u32* s_qn, u32& qsel)
{
    __syncthreads();
    const u32 qn = s_qn[qsel];
    if (qn) {
        // entries past the capacity were tested by the lanes that met them (qcf_filter_spill)
        qcf_filter_drain<LN, LG>(prm, s_q[qsel], (qn < QCF_FILTER_QCAP) ? qn : QCF_FILTER_QCAP);
        if (prm.dbg && (threadIdx.x == 0)) {
            atomicAdd(&prm.dbg[0], qn);
            atomicMax(&prm.dbg[1], qn);
        }
    }
    __syncthreads();
    if (qn && (threadIdx.x == 0)) {
        s_qn[qsel] = 0u;
    }
    qsel ^= 1u;
}

// 128 threads = 4 warps = one per SM sub-partition (warps of a block are dealt to sub-partitions by warp id, so a
// 5-warp block loads one sub-partition twice and caps residency at 5 blocks -- measured 26 of 40 warps resident);
// 6 blocks = 24 warps per SM at <= 80 registers: measured 4 % faster than 8 blocks at 64 registers, where the loop
// counters of the segment and epoch level spill to local memory (the trip loop itself holds 32 offsets + ~12 values).
// Trips per flag test of the packed kernels (degree 1 / degree 2; at degree 2 a divisor of the 4 trips of a group).
// Measured on the B200 (tools/gpu_ab.sh, gpurun_out/r2t..r2v_ab.txt, bench launch shapes, ms per launch):
//   degree 2, 96 bit:   1 trip per test 6.10,  2: 5.86,  4: 5.89
//   degree 1, 128 bit:  1: 37.5,  2: 34.9,  4: 33.4,  8: 32.6;   112 bit (8 segments of 28 trips):  1: 40.8,  2: 38.4,  4: 37.3,  8: 39.0
// A warp-wide OR reduction into a uniform register (QCF_FILTER_REDUX) instead of the vote: 6.21 at degree 2, 33.2 at degree 1.
// Blocks per SM at degree 2: 5 (96 registers) 5.83, 6 (80) 5.86-5.94, 7 (72) 6.51.  The group moves EG[] in shared memory
// instead of 8 registers: 5.96 (and 2 % slower down the last slice of an 8-way split).
#ifndef QCF_FILTER_CHECK1
#define QCF_FILTER_CHECK1 4u
#endif
#ifndef QCF_FILTER_CHECK2
#define QCF_FILTER_CHECK2 2u
#endif
// "some lane of the warp raised its flag": a vote on a per-lane compare, or (QCF_FILTER_REDUX) one warp-wide OR reduction
// whose result lands in a uniform register -- the compare and the branch then run on the uniform datapath
#ifndef QCF_FILTER_REDUX
#define QCF_FILTER_REDUX 0
#endif
__device__ __forceinline__ bool qcf_warp_any_nonzero(u32 x)
{
#if defined(__CUDA_ARCH__) && QCF_FILTER_REDUX
    return __reduce_or_sync(0xffffffffu, x) != 0u;
#else
    return __any_sync(0xffffffffu, x != 0u);
#endif
}
#define QCF_FILTER_BLOCK 128
#define QCF_FILTER_ROWS 4u
#ifndef QCF_FILTER_MIN_BLOCKS
#define QCF_FILTER_MIN_BLOCKS 6
#endif

// Every chain of a launch has exactly K = prm.steps guesses and walks whole trips (the last one padded past K).
// A row = (c, j) = n_a chains.  All threads of a block walk their chains in step (same trip counts), which is what lets
// the block meet at the end of every epoch to drain the survivor queue.
// PK: the packed 16-bit form of the trip (qcf_filter_math.cuh): always at degree 1; at degree 2 when the host found the trip
// moves exact on the high halves (FilterPlan::packed); never at degree 3.
template <int D, int LN, int LG, bool PK> __global__ void __launch_bounds__(QCF_FILTER_BLOCK, QCF_FILTER_MIN_BLOCKS) qcf_filter_kernel(const __grid_constant__ FilterParams prm)
{
    extern __shared__ u32 s_tab[];
    u32* s_a = s_tab;
    u32* s_b = s_tab + prm.n_a;
    __shared__ uint4 s_q[2][QCF_FILTER_QCAP];
    __shared__ u32 s_qn[2];
    __shared__ uint4 s_seg[QCF_FILTER_MAX_SEG]; // (trips, shift, bound, slope change high limb) of each segment
    __shared__ u32 s_nulo[QCF_FILTER_MAX_SEG];  // slope change, low limb
    __shared__ u32 s_kappa[QCF_FILTER_MAX_SEG]; // change of the per-chain offset correction
    __shared__ u32 s_rho[QCF_FILTER_BLOCK];     // the lane's chain offset (read once per segment: not a register of the chain loop)
    __shared__ u64 s_claim;
    // what a lane needs only when it queues a survivor lives in shared memory, not in registers of the chain loop:
    // (chain offset c, chain residue g0s | idle flag)
    __shared__ uint4 s_chain[QCF_FILTER_BLOCK];
    for (u32 i = threadIdx.x; i < prm.n_a; i += blockDim.x) {
        s_a[i] = prm.d_a[i];
    }
    for (u32 i = threadIdx.x; i < prm.n_b; i += blockDim.x) {
        s_b[i] = prm.d_b[i];
    }
    if (threadIdx.x < prm.n_seg) {
        s_seg[threadIdx.x] = make_uint4(prm.seg_trips[threadIdx.x], prm.seg_shift[threadIdx.x], prm.seg_bound[threadIdx.x], prm.seg_nu_hi[threadIdx.x]);
        s_nulo[threadIdx.x] = prm.seg_nu_lo[threadIdx.x];
        s_kappa[threadIdx.x] = prm.seg_kappa[threadIdx.x];
    }
    if (threadIdx.x < 2u) {
        s_qn[threadIdx.x] = 0u;
    }
    __syncthreads();

    const u32 P = prm.period, n_a = prm.n_a, n_b = prm.n_b, tp = prm.tprime, al = prm.align;
    const u32 kc = prm.steps, n_seg = prm.n_seg, epoch = prm.epoch_trips;
    const bool drift = prm.drift != 0u;
    const u64 base64 = ((u64)prm.base[1] << 32) | prm.base[0];
    const u64 n64 = prm.n64;
    constexpr u32 rows = QCF_FILTER_ROWS; // rows per claim: 4 * n_a chains = a whole number of chains per thread (n_a = 480 or 5760)
    const u64 n_srows = (prm.n_rows + rows - 1u) / rows;
    const u32 n_chains = rows * n_a;     // per row group; threads past the end walk along idle (barriers)
    constexpr u32 U = QCF_FILTER_UNROLL_FOR(D);

    // Row groups are claimed from a device counter, not strided: warps of co-resident blocks progress at different
    // rates (the scheduler favours some), and with static shares the SM idles while the slowest block finishes.
    u32 qsel = 0;       // the queue this epoch appends to
    u32 budget = epoch; // trips left in this epoch; epochs run across segments, chains and row groups
    u64 tested = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            // the found flag (finish() on another GPU/thread) ends the launch at the next claim, uniformly for the block
            u64 claim = (*(volatile u32*)&prm.hits->stop) ? ~0ull : atomicAdd(prm.work, 1ull);
            if (prm.rand_on) {
                // chain-draw random mode: claim number j is the j-th SAMPLE of the launch; Philox draws its row group
                claim = (claim < prm.rand_count)
                    ? qcf_random_row_group_of(prm.rand_key0, prm.rand_key1, prm.rand_block_lo, prm.rand_block_hi, prm.rand_first + (u32)claim, n_srows)
                    : ~0ull;
            }
            s_claim = claim;
        }
        __syncthreads();
        const u64 srow = s_claim;
        if (srow >= n_srows) {
            break;
        }
        // every thread makes the same number of rounds (4 * n_a is a multiple of the block size; n_a alone is not)
        for (u32 e0 = 0; e0 < n_chains; e0 += blockDim.x) {
            const u32 e = e0 + threadIdx.x;
            const u32 rr = (e < n_chains) ? e / n_a : 0u;
            const u32 i = (e < n_chains) ? e - rr * n_a : 0u;
            const u64 row = srow * rows + rr;
            const bool active = (e < n_chains) && (row < prm.n_rows); // idle threads walk along (barriers) but queue nothing
            u64 c = row;
            u32 bj = 0;
            if (n_b > 1u) {
                c = __umul64hi(row, prm.nb_magic); // row / n_b, exact for row < 2^64 / n_b
                bj = s_b[(u32)(row - c * n_b)];
            }
            u32 g0s = s_a[i] + bj;
            g0s = (g0s >= P) ? (g0s - P) : g0s;
            const u64 g0 = base64 + c * P + g0s;
            s_chain[threadIdx.x] = make_uint4((u32)c, (u32)(c >> 32), g0s, active ? kc : 0u);
            // ---- chain setup: x = g0^-1 mod 2^64, then the polynomial coefficients of Z in k (qcf_filter_math.cuh) ----
            u32 z, d1, elo, d2, d3;
            {
                const u32 rho32 = prm.rho ? qcf_filter_rho32(c, g0s, prm.pinv32, tp) : 0u;
                s_rho[threadIdx.x] = rho32;
                qcf_filter_chain_init<D>(n64, g0, P, tp, al, prm.c_lo, prm.mu0, prm.rho ? qcf_filter_rho_corr(rho32, prm.sigma0) : 0ull, prm.slack, z, d1, elo, d2, d3);
            }
            // ---- walk the chain, U guesses per trip ----
            // Inside a trip starting at step k0 the tracked limb is  z(k0 + u) = z + s[u]  with the trip-relative
            // offsets  s[u] = u*d1 + C(u,2)*d2 + C(u,3)*d3  (d1, d2 = differences at k0).  One fused add+min per
            // guess tests the trip; moving to the next trip costs D-1 multiply-adds per guess on the other pipe:
            //   D=1: s[u] constant;  D=2: s[u] += u*(U*d2);  D=3: s[u] += u*E + C(u,2)*F, E = U*d2 + C(U,2)*d3, F = U*d3.
            if constexpr (D == 1) {
                // ---- degree 1: the packed 16-bit trip (qcf_filter_math.cuh): two guesses per add+min, the high halves only;
                // a trip that may hold a survivor is replayed with the full 32-bit comparison ----
                u32 S2[U / 2];
                qcf_filter_pack_offsets<U>(S2, d1);
                u32 k = 0;
                for (u32 sg = 0; sg < n_seg; ++sg) {
                    const uint4 seg = s_seg[sg];
                    if (drift && (sg > 0u)) { // re-base onto this segment's reference line: value and slope
                        qcf_filter_enter_segment1(z, d1, elo, seg.y, s_nulo[sg], seg.w, s_rho[threadIdx.x], s_kappa[sg]);
                        qcf_filter_pack_offsets<U>(S2, d1);
                    } else {
                        z += seg.y; // re-base onto this segment's cofactor window
                    }
                    const u32 bound = seg.z;
                    const u32 t2 = ((bound >> 16) + 2u) * 0x10001u; // T in both halves (the planner keeps bound < 2^26)
                    u32 m0 = t2, m1 = t2;
                    for (u32 left = seg.x; left > 0u;) {
                        const u32 run = (left < budget) ? left : budget;
                        left -= run;
                        budget -= run;
                        // C trips under ONE flag test (see degree 2 below); rare: replay them at full width and queue their
                        // survivors -- the whole warp together (a lane that replayed alone never rejoined its warp: the replay holds
                        // a loop, and ncu showed 4 active lanes per issue).  Most flags are the packed test's roundings: the
                        // full-width minimum over the quarter of the steps the flag names (u mod 4) decides first.
                        const auto trips = [&](auto cnt) {
                            constexpr u32 C = decltype(cnt)::value;
                            const u32 zt = z;
#pragma unroll
                            for (u32 ct = 0; ct < C; ++ct) {
                                qcf_filter_trip16<U>(z, S2, m0, m1, t2);
                                z += U * d1; // to the next trip
                            }
                            if (qcf_warp_any_nonzero((m0 ^ t2) | (m1 ^ t2))) {
                                const u32 cls = qcf_filter_flag_classes(m0, m1, t2);
                                m0 = m1 = t2;
                                if (prm.dbg) {
                                    atomicAdd(&prm.dbg[2], 1u);
                                }
                                const u32 mn = qcf_filter_flagged_min<C * U>(zt, d1, 0u, cls);
                                if (__any_sync(0xffffffffu, mn <= bound)) {
                                    qcf_filter_queue_classes<LN, LG, C * U>(prm, s_chain[threadIdx.x], s_q[qsel], &s_qn[qsel], zt, d1, 0u, cls, bound, k);
                                }
                            }
                            k += C * U;
                        };
                        u32 r = run;
                        for (; r >= QCF_FILTER_CHECK1; r -= QCF_FILTER_CHECK1) {
                            trips(std::integral_constant<u32, QCF_FILTER_CHECK1>());
                        }
                        for (; r > 0u; --r) {
                            trips(std::integral_constant<u32, 1u>());
                        }
                        if (budget == 0u) {
                            qcf_filter_epoch_end<LN, LG>(prm, s_q, s_qn, qsel);
                            budget = epoch;
                        }
                    }
                }
            } else if constexpr ((D == 2) && PK) {
                // ---- degree 2, packed: sub-trips of 16 guesses in groups of 8 (qcf_filter_math.cuh): the offsets of sub-trip j
                // come from the group's exact base by one multiply-add per register, the base moves by one packed add per
                // group; a trip (two sub-trips) that may hold a survivor is replayed at full width ----
                static_assert(U == 2u * QCF_FILTER_SUB, "a trip is two sub-trips");
                constexpr u32 TG = QCF_FILTER_GROUP / 2u; // trips per group
                u32 S2[QCF_FILTER_SUB / 2], E2[QCF_FILTER_SUB / 2], EG[QCF_FILTER_SUB / 2];
                qcf_filter_pack2_moves(E2, EG, d2);
                qcf_filter_pack2_base(S2, E2, d1, d2);
                u32 k = 0;
                u32 a = qcf_filter_sub_move(d1, d2); // the limb's move over one sub-trip; d1 at step k is d1z + k * d2
                u32 d1z = d1;
                for (u32 sg = 0; sg < n_seg; ++sg) {
                    const uint4 seg = s_seg[sg];
                    if (drift && (sg > 0u)) { // re-base onto this segment's reference line: value and slope; the base follows the slope
                        const u32 d1k = d1z + k * d2;
                        u32 d1n = d1k;
                        qcf_filter_enter_segment1(z, d1n, elo, seg.y, s_nulo[sg], seg.w, s_rho[threadIdx.x], s_kappa[sg]);
                        d1z += d1n - d1k;
                        a += QCF_FILTER_SUB * (d1n - d1k);
                        qcf_filter_pack2_base(S2, E2, d1n, d2);
                    } else {
                        z += seg.y;
                    }
                    const u32 bound = seg.z;
                    const u32 t2 = qcf_filter_t2_group(bound);
                    u32 m0 = t2, m1 = t2;
                    // C trips (sub-trips j .. j + 2C - 1 of the current group) under ONE flag test: the test -- two compares, a vote
                    // and a branch the next trip cannot start before -- is where the warps of this kernel wait (a third of the
                    // loop's stall samples with one test per trip); flagged -> replay the C * 32 steps from (zt, dt)
                    const auto trips = [&](const u32 j, auto cnt) {
                        constexpr u32 C = decltype(cnt)::value;
                        const u32 zt = z;
#pragma unroll
                        for (u32 st = 0; st < 2u * C; ++st) {
                            qcf_filter_subtrip16(z, j + st, S2, E2, m0, m1);
                            qcf_filter_sub_advance(z, a, d2);
                        }
                        if (qcf_warp_any_nonzero((m0 ^ t2) | (m1 ^ t2))) { // rare; the whole warp replays together
                            const u32 cls = qcf_filter_flag_classes(m0, m1, t2);
                            const u32 dt = d1z + k * d2; // the first difference at the first step
                            m0 = m1 = t2;
                            if (prm.dbg) {
                                atomicAdd(&prm.dbg[2], 1u);
                            }
                            // most flags are the packed test's roundings: the full-width minimum decides first -- over the quarter
                            // of the steps the flag names (u mod 4), evaluated directly
                            const u32 mn = qcf_filter_flagged_min<C * U>(zt, dt, d2, cls);
                            if (__any_sync(0xffffffffu, mn <= bound)) {
                                qcf_filter_queue_classes<LN, LG, C * U>(prm, s_chain[threadIdx.x], s_q[qsel], &s_qn[qsel], zt, dt, d2, cls, bound, k);
                            }
                        }
                        k += C * U;
                    };
                    const auto trip = [&](const u32 j) { trips(j, std::integral_constant<u32, 1u>()); };
                    for (u32 left = seg.x; left > 0u;) {
                        const u32 run = (left < budget) ? left : budget;
                        left -= run;
                        budget -= run;
                        u32 r = run;
                        for (; r >= TG; r -= TG) { // whole groups
#pragma unroll
                            for (u32 t = 0; t < TG; t += QCF_FILTER_CHECK2) {
                                trips(2u * t, std::integral_constant<u32, QCF_FILTER_CHECK2>());
                            }
#pragma unroll
                            for (u32 p = 0; p < QCF_FILTER_SUB / 2; ++p) {
                                S2[p] = qcf_add16x2(S2[p], EG[p]);
                            }
                        }
                        if (r) { // the rest of the run, then the base moves on exactly: one packed add per sub-trip walked
                            u32 j = 0;
                            for (; r > 0u; --r, j += 2u) {
                                trip(j);
                            }
                            if ((left > 0u) || !drift) { // at the end of a segment the next one rebuilds the base (drift lines)
                                for (; j > 0u; --j) {
                                    qcf_filter_pack2_step_base(S2, E2);
                                }
                            }
                        }
                        if (budget == 0u) {
                            qcf_filter_epoch_end<LN, LG>(prm, s_q, s_qn, qsel);
                            budget = epoch;
                        }
                    }
                }
            } else {
                u32 s[U];
                qcf_filter_offsets<U>(s, d1, d2, d3);
                const u32 f3 = U * d3;
                u32 k = 0;
                for (u32 sg = 0; sg < n_seg; ++sg) {
                    const uint4 seg = s_seg[sg];
                    if (drift && (sg > 0u)) { // re-base onto this segment's reference line: value and slope
                        qcf_filter_enter_segment<U>(z, d1, elo, s, seg.y, s_nulo[sg], seg.w, s_rho[threadIdx.x], s_kappa[sg]);
                    } else {
                        z += seg.y; // re-base onto this segment's cofactor window
                    }
                    const u32 bound = seg.z;
                    for (u32 left = seg.x; left > 0u;) {
                        const u32 run = (left < budget) ? left : budget;
                        left -= run;
                        budget -= run;
                        for (const u32 k_end = k + run * U; k < k_end; k += U) {
                            u32 mn = z, mn2 = z + s[1];
    #pragma unroll
                            for (u32 u = 2; u < U; u += 2) {
                                mn = min(mn, z + s[u]);
                                mn2 = min(mn2, z + s[u + 1]);
                            }
                            mn = min(mn, mn2);
                            if (mn <= bound) { // rare: queue the survivors of this trip
    #pragma unroll
                                for (u32 u = 0; u < U; ++u) {
                                    if (z + s[u] <= bound) {
                                        qcf_filter_queue<LN, LG>(prm, s_chain[threadIdx.x], s_q[qsel], &s_qn[qsel], k + u);
                                    }
                                }
                            }
                            qcf_filter_advance<D, U>(z, d1, d2, d3, f3, s); // to the next trip
                        }
                        if (budget == 0u) {
                            qcf_filter_epoch_end<LN, LG>(prm, s_q, s_qn, qsel);
                            budget = epoch;
                        }
                    }
                }
            }
            if (active && prm.count) {
                const uint4 ch = s_chain[threadIdx.x];
                tested += prm.n_extra ? qcf_chain_count_extra(prm, ((u64)ch.y << 32) | ch.x, ch.z) : kc;
            }
        }
    }
    qcf_filter_epoch_end<LN, LG>(prm, s_q, s_qn, qsel); // what the last, unfinished epoch queued
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

// `grid` comes in as the number of SMs; the launch uses exactly as many blocks as are resident at once (one wave:
// a partial second wave would run at a fraction of the occupancy), capped by the work available.
template <int D, int LN, int LG, bool PK> static cudaError_t qcf_launch_filter_t(const FilterParams& prm, int grid, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
    // attribute once per device, occupancy once per (device, shared-memory size): qcf_resident_blocks (qcf_internal.h)
    static QcfLaunchCache cache;
    int per_sm = 0;
    const cudaError_t e = qcf_resident_blocks(qcf_filter_kernel<D, LN, LG, PK>, cache, QCF_FILTER_BLOCK, smem, &per_sm);
    if (e != cudaSuccess) {
        return e;
    }
    const u64 srows = prm.rand_on ? (u64)prm.rand_count : (prm.n_rows + QCF_FILTER_ROWS - 1u) / QCF_FILTER_ROWS;
    const u64 blocks = (u64)grid * (u64)(per_sm > 0 ? per_sm : 1);
    qcf_filter_kernel<D, LN, LG, PK><<<(unsigned)(blocks < srows ? blocks : srows), QCF_FILTER_BLOCK, smem, stream>>>(prm);
    return cudaGetLastError();
}

cudaError_t qcf_launch_filter(const FilterParams& prm, int degree, bool packed, int ln, int lg, int grid, cudaStream_t stream)
{
#define QCF_CASE(LN, LG)                                                                                               \
    if ((ln == LN) && (lg == LG)) {                                                                                    \
        if (degree == 1) {                                                                                             \
            return qcf_launch_filter_t<1, LN, LG, true>(prm, grid, stream);                                            \
        }                                                                                                              \
        if (degree == 2) {                                                                                             \
            return packed ? qcf_launch_filter_t<2, LN, LG, true>(prm, grid, stream)                                    \
                          : qcf_launch_filter_t<2, LN, LG, false>(prm, grid, stream);                                  \
        }                                                                                                              \
        if (degree == 3) {                                                                                             \
            return qcf_launch_filter_t<3, LN, LG, false>(prm, grid, stream);                                           \
        }                                                                                                              \
    }
    QCF_CASE(2, 1) QCF_CASE(2, 2) QCF_CASE(3, 1) QCF_CASE(3, 2) QCF_CASE(4, 1) QCF_CASE(4, 2) QCF_CASE(4, 3)
#undef QCF_CASE
    return QCF_NO_KERNEL;
}
