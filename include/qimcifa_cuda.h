/*
 * qimcifa_cuda.h -- C ABI of libqimcifa_cuda.so: the B200 (sm_100a) implementation of qimcifa's
 * reverse-wheel trial-division sweep.
 *
 * The reference has no plugin/FFI interface; its hot path is the inline header function
 *     Qimcifa::getSmoothNumbers(toFactor, inc_seqs, offset, iterClock)   reference include/qimcifa.hpp:384-399
 * plus the dispenser globals it reads (include/qimcifa.hpp:97-129).  This header is the narrowest
 * cut at which everything above stays host C++ and everything below runs on the GPU.  Every entry
 * point names the reference interface it replaces.  Plain C types, caller-owned buffers, no C++
 * and no torch types across the boundary.  All functions return 0 on success and a negative
 * QCF_E* code on error (the reference checks nothing); they never throw and
 * never abort.  qcf_last_error() returns the message of the last failure on the calling thread.
 *
 * Conventions
 *   - Big integers are little-endian arrays of 32-bit limbs (n_le[0] = least significant).
 *   - "index" = 0-based rank p among integers coprime to 2310; forward(p) is its value
 *     (include/qimcifa.hpp:256-259).  "batch b" owns indices [b*BW+2, (b+1)*BW+1],
 *     BW = 223092870 (include/qimcifa.hpp:86,388-392).
 *   - "level" L in [5,10] = number of wheel primes of {2,3,5,7,11,13,17,19,23,29}
 *     (src/qimcifa.cpp:64,93-98).  A guess is tested iff it is coprime to all of them.  Level 10 runs on the
 *     level-9 tables and drops the multiples of 29 (its own period does not fit 32 bits); the random-guess
 *     mode and qcf_get_tables take levels 5..9 only.
 *   - One context per GPU; a context is not thread-safe, different contexts are independent.
 */
#ifndef QIMCIFA_CUDA_H
#define QIMCIFA_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCF_BIGGEST_WHEEL 223092870ULL /* include/qimcifa.hpp:86 */
#define QCF_MIN_LEVEL 5                /* include/qimcifa.hpp:89 */
#define QCF_MAX_LEVEL 10               /* src/qimcifa.cpp:64 lists 29 and the prompt says "max of 10"; the reference then
                                        * clamps to 9 (:96-98), and so do the host drivers here unless --level-10 */
#define QCF_MAX_N_LIMBS 4              /* N < 2^128 */
#define QCF_MAX_HITS 64

#define QCF_OK 0
#define QCF_EINVAL (-1)  /* bad argument */
#define QCF_ECUDA (-2)   /* CUDA runtime error */
#define QCF_ERANGE (-3)  /* value does not fit the supported widths */
#define QCF_ENODEV (-4)  /* no CUDA device: there is NO CPU fallback */

/* kernel selection for qcf_sweep*(..., flags) */
#define QCF_KERNEL_AUTO 0u   /* filter kernel where it applies, exact kernel elsewhere */
#define QCF_KERNEL_EXACT 1u  /* one exact Hensel/Montgomery divisibility test per guess */
#define QCF_KERNEL_FILTER 2u /* 2-adic quotient filter along arithmetic progressions + exact test of survivors */
#define QCF_KERNEL_MASK 3u
#define QCF_COUNT_ON_DEVICE 4u /* also count tested guesses on the device (cross-check of the closed form) */
/* Common-factor mode: what the reference does when built with IS_RSA_SEMIPRIME=OFF (CMakeLists.txt:7,
 * include/qimcifa.hpp:201-210,373-379): a guess is reported when gcd(guess, N) != 1, and factor_le receives that
 * gcd ("Has common factor: Found gcd * N/gcd"); qcf_last_hits returns the guesses themselves. */
#define QCF_MODE_GCD 8u

typedef struct qcf_ctx qcf_ctx;

typedef struct qcf_result {
    int32_t found;            /* 1 if a divisor of N was found (printSuccess, include/qimcifa.hpp:212-222) */
    int32_t aborted;          /* 1 if the sweep stopped early on the found flag (finish(), include/qimcifa.hpp:102-105) */
    uint32_t factor_le[QCF_MAX_N_LIMBS]; /* the guess that divides N; host computes N / factor as the reference does (:370) */
    uint32_t n_hits;          /* divisors seen in the launch that ended the sweep */
    uint32_t kernel_kind;     /* QCF_KERNEL_EXACT or QCF_KERNEL_FILTER: what the last launch ran */
    uint64_t guesses_tested;  /* exact number of wheel-filtered guesses covered by completed launches (closed form) */
    uint64_t guesses_counted; /* device-side count when QCF_COUNT_ON_DEVICE, else 0 */
    uint64_t batches_done;    /* whole batches covered by completed launches */
    uint64_t next_batch_hi;   /* resume token: batches [batch_lo, next_batch_hi) are still unswept */
    uint64_t kernel_launches; /* sweep kernels launched by this call */
    double device_seconds;    /* CUDA-event time of those kernels on the context's stream */
} qcf_result;

/* ---- context ------------------------------------------------------------------------------- */
int qcf_create(int device, qcf_ctx** out);
int qcf_destroy(qcf_ctx* ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream) so the caller's CUDA events see the kernels. */
int qcf_set_stream(qcf_ctx* ctx, void* cuda_stream);
const char* qcf_last_error(void);
int qcf_device_info(qcf_ctx* ctx, int* sm_count, int* clock_khz, char* name, int name_len);

/* ---- wheel tables: replaces wheel_gen()/wheel_inc() (include/qimcifa.hpp:274-308) ----------- *
 * Generated ON THE DEVICE by a sieve + compaction kernel; cached per level in the context.
 * qcf_get_tables copies them out for tests: table A = residues coprime to the first min(L,6) primes
 * (pre-multiplied by the CRT coefficient at L>=7), table B = the CRT factor for primes 17..p_L. */
int qcf_build_tables(qcf_ctx* ctx, int level);
int qcf_get_tables(qcf_ctx* ctx, int level, uint32_t* period, uint32_t* n_a, uint32_t* n_b, uint32_t* table_a,
    uint32_t cap_a, uint32_t* table_b, uint32_t cap_b);

/* ---- the sweep: replaces getSmoothNumbers() (include/qimcifa.hpp:384-399) -------------------- *
 * Sweeps batches [batch_lo, batch_hi) in DESCENDING order, batches_per_launch batches per kernel
 * launch (0 = library default).  Stops after the first launch in which a divisor is found and
 * reports the divisor the reference's single-thread visiting order would report first (highest
 * batch, smallest guess inside it).  The found flag (qcf_set_found) is polled inside the kernel. */
int qcf_sweep(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t batch_lo, uint64_t batch_hi,
    uint64_t batches_per_launch, uint32_t flags, qcf_result* out);

/* Same, over an arbitrary inclusive index interval [p_lo, p_hi] (128-bit, 4 limbs each) in ONE launch
 * sequence; used by the parity tests and by ragged slice ends. */
int qcf_sweep_indices(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, const uint32_t* p_lo_le,
    const uint32_t* p_hi_le, uint32_t flags, qcf_result* out);

/* All divisors seen by the last sweep launch (up to QCF_MAX_HITS), 4 limbs each, unordered. */
int qcf_last_hits(qcf_ctx* ctx, uint32_t* hits_le, uint32_t cap, uint32_t* n_hits);

/* ---- range / slice arithmetic the host drivers need (src/qimcifa.cpp:128,149,155-158) -------- */
/* floor(sqrt(N)) (include/qimcifa.hpp:162-186), backward(floor sqrt) and the node split. */
int qcf_node_split(const uint32_t* n_le, int n_limbs, uint64_t node_count, uint32_t* sqrt_le /*4*/,
    uint32_t* full_range_le /*4*/, uint64_t* node_range, uint64_t* batch_count);
/* Exact number of guesses the level tests in batches [batch_lo, batch_hi). */
int qcf_count_guesses(int level, uint64_t batch_lo, uint64_t batch_hi, uint32_t* count_le /*4*/);

/* ---- found flag: replaces finish()/batchNumber=batchBound (include/qimcifa.hpp:102-105) ------ */
int qcf_set_found(qcf_ctx* ctx, int value); /* asynchronous: seen by a running sweep kernel */
int qcf_poll_found(qcf_ctx* ctx, int* value);

/* Cross-GPU found flag over peer memory (NVLink / NVSwitch), for one-process-per-GPU runs: replaces the human who
 * relays the answer between nodes (reference README.md:27) and finish() between threads.  Every rank exports a handle
 * of its found-flag word (CUDA IPC), all ranks exchange the handles out of band (64 bytes each, e.g. one all_gather)
 * and connect; from then on a sweep that finds a divisor writes the flag of every peer with a system-scope atomic
 * issued from a kernel on this GPU, and the peers' running kernels stop at their next poll -- no collective, no host
 * round trip.  qcf_peer_signal does the same on demand. */
#define QCF_PEER_HANDLE_BYTES 64
int qcf_peer_export(qcf_ctx* ctx, uint8_t* handle /* QCF_PEER_HANDLE_BYTES */);
int qcf_peer_connect(qcf_ctx* ctx, const uint8_t* handles /* n_ranks x QCF_PEER_HANDLE_BYTES */, int n_ranks, int self_rank);
int qcf_peer_signal(qcf_ctx* ctx);
int qcf_peer_disconnect(qcf_ctx* ctx);
/* Shared dispenser over peer memory: the reference's THREAD model -- every worker pulls the next run of batches from ONE
 * mutex-guarded descending counter (getNextBatch, include/qimcifa.hpp:107-129; src/qimcifa.cpp:160-178) -- carried across
 * GPUs and processes.  The counter is a word in the context of rank `owner_rank`; qcf_peer_dispense adds `take` to it
 * with a system-scope atomic issued from a kernel on the calling GPU (over NVLink when the owner is a peer) and returns
 * the value before the add: the caller then owns batches [total - start - take, total - start) of the node's slice.
 * All GPUs work at the top of the range, which is where a balanced semiprime's factor is. */
int qcf_dispenser_reset(qcf_ctx* ctx, uint64_t value);
int qcf_peer_dispense(qcf_ctx* ctx, int owner_rank, uint64_t take, uint64_t* start);

/* ---- tuner: replaces the timing loop of qimcifa_tuner (src/qimcifa_tuner.cpp:61-77,137-144) -- *
 * Times `batches` batches just below sqrt(N) at `level`; returns device seconds and exact guesses. */
int qcf_time_level(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t batches, uint32_t flags,
    double* seconds, uint64_t* guesses);

/* ---- random-guess Monte-Carlo mode ------------------------------------------------------------- *
 * The reference only describes this mode (README.md:25; IS_RANDOM is a dead macro, include/common/config.h.in:6):
 * there is no reference code, so only validity and determinism are checkable.  Guess number c (a 64-bit
 * counter) is drawn with the counter-based generator Philox4x32-10, key = seed, counter = (c, 0): words
 * (r0, r1, r2, r3) -> period m = mulhi64(r1:r0, M), A-index = mulhi32(r2, n_a), B-index = mulhi32(r3, n_b),
 * guess = m*P + (A[ia] + B[ib]) mod P, i.e. uniform over the level's guesses in the M = floor((sqrt(N)+1)/P)
 * whole wheel periods below sqrt(N); the value 1 is skipped.  Each guess takes the exact divisibility test.
 * Tests `budget` guesses with counters [counter_lo, counter_lo + budget); ranks of a multi-GPU run take
 * disjoint counter blocks.  The run does not stop at the first hit unless the found flag is raised. */
int qcf_sweep_random(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t budget, uint32_t flags, qcf_result* out);
/* Test hook: the guesses drawn for counters [counter_lo, counter_lo + count), 4 limbs each. */
int qcf_random_guesses(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t count, uint32_t* guesses_le);

/* Chain-draw random mode: qcf_sweep_random(..., flags = QCF_RANDOM_CHAINS).  The same generator, but a draw selects a
 * whole ROW GROUP of arithmetic progressions ("chains") of a random window of the range, and the filter kernel of the sweep
 * walks every guess of those chains: clustered Monte-Carlo sampling at the sweep's speed (~10^13 guesses/s on a B200 against
 * ~10^11 for independent draws, each of which pays a generator call, a table lookup and a full exact test).  Guess numbers
 * ("counters") map to guesses as follows; everything is a pure function of (N, level, seed, counter) for a given library:
 *   - counter block b = counter >> 34.  Window of the block: w = floor(x * n_win / 2^64), x = the first 64 bits of
 *     Philox4x32-10(key = seed, counter = (b_lo, b_hi, 0x51434657, 0)); window w covers the batches
 *     [max(0, hi - 4096), hi), hi = batchCount - 4096 w, n_win = ceil(batchCount / 4096) (batchCount of nodeCount = 1).
 *     Of a window's value range [v_lo, v_hi] only the top octave, above v_hi / 2, is drawn (the windows of a large number
 *     span far less than an octave: at 96 bits the lowest of 65 windows is the only one cut).
 *   - the library's filter plan for the window's whole wheel periods (qcf_last_plan returns it after the call) fixes the
 *     chain geometry: period P (of the chains' wheel level, <= the requested one), stride 2^tprime periods, K = steps,
 *     n_rows = n_b << tprime rows of n_a chains each; a row group = 4 rows; q = 4 n_a K guess numbers per row-group sample,
 *     R = floor(2^34 / q) samples per block (the last 2^34 - R q counters of a block are skipped).
 *   - counter c_in = counter mod 2^34: sample j = c_in / q, chain e = (c_in mod q) / K of the group, step k = c_in mod K.
 *     Row group of sample j: s = floor(x * ceil(n_rows / 4) / 2^64), x from Philox4x32-10(key = seed, counter = (b_lo, b_hi, j, 1)).
 *     row = 4 s + e / n_a, i = e mod n_a, (c, jb) = divmod(row, n_b):   guess = (m_lo + c + (k << tprime)) P + (A[i] + B[jb]) mod P.
 *     Guesses that are multiples of a wheel prime of the requested level are skipped (neither tested nor counted).
 * A call tests the samples whose FIRST counter lies in [counter_lo, counter_lo + budget): disjoint counter ranges test
 * disjoint sample sets.  Needs a number large enough for the filter kernel, and sqrt(N) above half the top of the padded range
 * (else QCF_ERANGE: use flags = 0, as `qimcifa --random` does by itself). */
#define QCF_RANDOM_CHAINS 16u
/* Test hook: the guesses of counters [counter_lo, counter_lo + count) under the chain-draw map (0 = skipped), 4 limbs each. */
int qcf_random_chain_guesses(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t count, uint32_t* guesses_le);

/* ---- limb-kernel test hooks ------------------------------------------------------------------ */
/* out[i] = (N % g_i == 0) for `count` odd guesses of g_limbs limbs each, through the same device
 * routine the sweep kernels use (the replacement of BigInteger operator%, include/qimcifa.hpp:369). */
int qcf_divisible(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, const uint32_t* g_le, int g_limbs, uint64_t count,
    uint8_t* out);

/* Integer-pipe micro-benchmark: issues `iters` x 32 independent 32-bit IMADs (kind 0), IMAD.WIDE.U32
 * (kind 1), IADD3 (kind 2) or IMAD.HI.U32 (kind 3) per thread on every SM; returns thread-level operations per second.  Used
 * to measure the roofline denominator of BASELINE.md section 3 on the box the bench runs on. */
int qcf_microbench(qcf_ctx* ctx, int kind, uint64_t iters, double* ops_per_second, double* seconds);

/* Planner test hook (host only, needs no GPU): the filter-kernel launch the library would make for the whole wheel
 * periods of the level inside the inclusive VALUE interval [v_lo, v_hi] (4 limbs each).  The CPU tests run the
 * kernel's arithmetic, restated in Python, on exactly these parameters.  found = 0: the filter does not apply. */
#define QCF_PLAN_MAX_SEG 32
typedef struct qcf_filter_plan {
    int32_t found, degree;
    uint32_t tprime, align, steps, trips, unroll, n_seg, slack, fbits;
    uint32_t seg_trips[QCF_PLAN_MAX_SEG], seg_shift[QCF_PLAN_MAX_SEG], seg_bound[QCF_PLAN_MAX_SEG];
    uint32_t m_lo_le[4], m_end_le[4]; /* periods [m_lo, m_end) are covered */
    uint64_t c_lo;    /* the reference line of segment 0 at step 0, mod 2^64 */
    uint32_t period;
    double cost;
    /* linear cofactor drift: inside segment j the quotient is compared against the line
     *   l_j(k) = seg_beta[j] - seg_lambda[j] * (k - seg_ka[j])        (k = step of the chain, counted from the launch's first)
     * which lies below the cofactor of every guess of the segment by at most seg_window[j]; drift = 0: seg_lambda = 0 (constant
     * floors).  mu0 = seg_lambda[0] << align; seg_nu[j] = (seg_lambda[j] - seg_lambda[j-1]) << align (64 bits, two limbs). */
    uint32_t drift;
    uint64_t mu0;
    uint32_t seg_nu_lo[QCF_PLAN_MAX_SEG], seg_nu_hi[QCF_PLAN_MAX_SEG];
    uint64_t seg_lambda[QCF_PLAN_MAX_SEG], seg_ka[QCF_PLAN_MAX_SEG], seg_window[QCF_PLAN_MAX_SEG];
    uint32_t seg_beta_le[QCF_PLAN_MAX_SEG][4];
    /* per-chain offset correction (rho = 1): a chain whose first guess lies rho above the launch's first value lowers its
     * reference by A_j = what qcf_filter_rho32 / qcf_filter_rho_corr (csrc/qcf_filter_math.cuh) compute from the bound
     * seg_sigma[j] on one step's cofactor drift (sigma0 = seg_sigma[0]; seg_kappa[j] = ((seg_sigma[j-1] - seg_sigma[j]) << align) >> 32) */
    uint32_t rho, pinv32;
    uint64_t sigma0;
    uint64_t seg_sigma[QCF_PLAN_MAX_SEG];
    uint32_t seg_kappa[QCF_PLAN_MAX_SEG];
    /* packed = 1: the kernel walks the trips in the packed 16-bit form (two guesses per add+min on the high halves of the
     * tracked limb, flagged trips replayed at full width): always at degree 1, at degree 2 when the stride allows it */
    uint32_t packed, reserved_;
} qcf_filter_plan;
int qcf_plan_filter(const uint32_t* n_le, int n_limbs, int level, const uint32_t* v_lo_le, const uint32_t* v_hi_le,
    int forced, qcf_filter_plan* out);

/* The plan of the largest filter-kernel launch (most periods) of the last qcf_sweep / qcf_sweep_indices / qcf_time_level call
 * on this context; found = 0 when that call launched no filter kernel.  The parity tests use it to prove which launch
 * shapes (degree, stride, segment count) a campaign exercised. */
int qcf_last_plan(qcf_ctx* ctx, qcf_filter_plan* out);

/* Survivors of the filter kernel are queued in shared memory and tested together at the end of an epoch; a survivor that
 * finds the queue full is tested on the spot by the lane that met it (nothing is ever dropped).  This returns how many
 * took that path in the last sweep call on this context (a statistic for tests and tuning; same per-guess test as
 * include/qimcifa.hpp:369 either way). */
int qcf_last_queue_spills(qcf_ctx* ctx, uint64_t* spills);

/* Bytes that cross the host/device boundary per sweep call, for end-to-end accounting: the kernel parameter block of a
 * chain launch (filter / exact chain / common-factor chain kernels) and of a row launch go up with the launch, the hit
 * record comes down once per call.  N itself is 16 bytes inside the parameter block. */
int qcf_io_bytes(uint32_t* h2d_per_chain_launch, uint32_t* h2d_per_row_launch, uint32_t* d2h_per_call);

/* ---- test / calibration switches --------------------------------------------------------------- *
 * Process-wide, validated, read at the entry of every sweep call; the library reads nothing from the environment.
 * value = NULL or "" restores the default.  None is needed in normal use.
 *   "xchain_two_stage"  0 | 1        exact chain kernel: full reduction per guess / low limb first (default 1)
 *   "chain_min_log2"    16..40       smallest run of periods (log2) that takes the chain form of the exact kernels (19)
 *   "plan_force"        "D,tprime,segments[,pad]"  restricts the filter planner's search (0 = free)
 *   "plan_drift"        -1 | 0 | 1 | 2  filter windows: 0 constant cofactor floors, 1 lines that follow the cofactor's drift,
 *                                    2 (= -1, the default) the same plus the per-chain offset correction
 *   "plan_packed"       0 | 1        degree-2 filter launches in the packed 16-bit form where the stride allows it (default 1)
 *   "cost_model"        8 numbers    replaces the planner's fitted cost constants (0 keeps one)
 *   "epoch_trips"       0..2^24      trips between two drains of the filter kernel's survivor queue (0 = automatic)
 *   "debug", "debug_stats"  0 | 1    plans / launches / survivor statistics on stderr
 * Returns QCF_EINVAL for an unknown name or a value outside the stated range. */
int qcf_set_option(const char* name, const char* value);

#ifdef __cplusplus
}
#endif
#endif /* QIMCIFA_CUDA_H */
