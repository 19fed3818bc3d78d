// C ABI of libqimcifa_cuda.so (see include/qimcifa_cuda.h): host-side range arithmetic, launch
// sequencing and result selection.  There is NO CPU fallback in this file: every sweep runs on the GPU.
//
// Host-side reference logic restated here (paths relative to the reference checkout):
//   forward / backward          include/qimcifa.hpp:256-263
//   floor sqrt                  include/qimcifa.hpp:162-186
//   range / node split          src/qimcifa.cpp:128,149,155-158
//   batch index bounds          include/qimcifa.hpp:388-392
//   descending batch order      include/qimcifa.hpp:119-127
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "qcf_internal.h"
#include "qcf_philox.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define QCF_CUDA(call)                                                                                                 \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) {                                                                                       \
            return fail(QCF_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);        \
        }                                                                                                              \
    } while (0)

// ---- test / calibration switches: qcf_set_option (nothing is read from the environment) ----------------------------------
std::mutex g_opt_mu;
QcfOptions g_opt;

bool parse_ints(const char* v, int* out, int min_n, int max_n, long lo, long hi)
{
    int n = 0;
    const char* p = v;
    while (*p) {
        char* end = nullptr;
        const long x = strtol(p, &end, 10);
        if ((end == p) || (x < lo) || (x > hi) || (n >= max_n)) {
            return false;
        }
        out[n++] = (int)x;
        p = end;
        if (*p == ',') {
            ++p;
            if (!*p) {
                return false;
            }
        } else if (*p) {
            return false;
        }
    }
    return n >= min_n;
}

const unsigned PRIMES10[10] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29 }; // src/qimcifa.cpp:64

// Level 10 (prime 29) has no tables of its own -- its period 2*3*...*29 = 6.5e9 does not fit the 32-bit residue
// tables: it runs on the level-9 (or lower, superset) tables and drops the multiples of 29 where guesses are counted and
// reported.
constexpr int QCF_TABLE_MAX_LEVEL = 9;
inline int table_level(int level) { return (level > QCF_TABLE_MAX_LEVEL) ? QCF_TABLE_MAX_LEVEL : level; }

struct Wheel11 {
    uint16_t w[480];
    Wheel11()
    {
        int k = 0;
        for (int i = 1; i < 2310; ++i) {
            if ((i % 2) && (i % 3) && (i % 5) && (i % 7) && (i % 11)) {
                w[k++] = (uint16_t)i;
            }
        }
    }
};
const Wheel11 W11;

inline u128 from_limbs(const u32* le, int n)
{
    u128 v = 0;
    for (int i = n - 1; i >= 0; --i) {
        v = (v << 32) | le[i];
    }
    return v;
}
inline void to_limbs(u128 v, u32* le, int n)
{
    for (int i = 0; i < n; ++i) {
        le[i] = (u32)v;
        v >>= 32;
    }
}
inline int bitlen(u128 v)
{
    const u64 hi = (u64)(v >> 64), lo = (u64)v;
    return hi ? (128 - __builtin_clzll(hi)) : (lo ? (64 - __builtin_clzll(lo)) : 0);
}

inline u128 h_forward(u128 p) { return (u128)W11.w[(size_t)(p % 480U)] + (p / 480U) * 2310U; }
inline u128 h_backward(u128 n)
{
    const unsigned r = (unsigned)(n % 2310U);
    const unsigned lb = (unsigned)(std::lower_bound(W11.w, W11.w + 480, (uint16_t)r) - W11.w);
    return (u128)lb + (u128)480U * (n / 2310U) + 1U;
}
inline u128 h_isqrt(u128 n)
{
    if (n < 2) {
        return n;
    }
    u128 x = (u128)1 << ((bitlen(n) + 1) / 2);
    for (;;) {
        const u128 y = (x + n / x) >> 1;
        if (y >= x) {
            return x;
        }
        x = y;
    }
}

// #{1 <= v <= x : v coprime to the first `level` primes}
u128 count_coprime_upto(u128 x, int level)
{
    if (x == 0) {
        return 0;
    }
    __int128 total = 0;
    for (unsigned mask = 0; mask < (1U << level); ++mask) {
        u128 d = 1;
        int bits = 0;
        for (int i = 0; i < level; ++i) {
            if ((mask >> i) & 1U) {
                d *= PRIMES10[i];
                ++bits;
            }
        }
        total += (bits & 1) ? -(__int128)(x / d) : (__int128)(x / d);
    }
    return (u128)total;
}

inline u128 batch_p_lo(u64 b) { return (u128)b * QCF_BIGGEST_WHEEL + 2U; }
inline u128 batch_p_hi(u64 b) { return ((u128)b + 1U) * QCF_BIGGEST_WHEEL + 1U; }
inline u128 batch_of_index(u128 p) { return (p < 2U) ? 0 : (p - 2U) / QCF_BIGGEST_WHEEL; }

} // namespace

namespace {
struct FilterPlan {
    int degree = 0;
    bool packed = false;      // the trip runs in the packed 16-bit form (qcf_filter_math.cuh): degree 1 always, degree 2 when its moves are exact on the high halves
    u32 tprime = 0, align = 0, bound = 0, slack = 0, fbits = 0, steps = 0, n_seg = 1;
    u32 trips = 0;            // whole trips a chain walks: trips * unroll >= steps
    u128 m_lo = 0, m_end = 0; // whole periods covered: [m_lo, m_end), m_end - m_lo = steps << tprime
    u64 c_lo = 0;
    u32 seg_trips[QCF_FILTER_MAX_SEG] = {};
    u32 seg_shift[QCF_FILTER_MAX_SEG] = {};
    u32 seg_bound[QCF_FILTER_MAX_SEG] = {};
    // linear cofactor drift (FilterParams::drift): the reference line of segment j is  l_j(k) = beta_j - lambda_j * (k - ka_j)
    bool drift = false;
    u64 mu0 = 0;                              // lambda_0 << align
    u32 seg_nu_lo[QCF_FILTER_MAX_SEG] = {};   // (lambda_j - lambda_{j-1}) << align, 64 bits
    u32 seg_nu_hi[QCF_FILTER_MAX_SEG] = {};
    u64 seg_lambda[QCF_FILTER_MAX_SEG] = {};  // for the tests: the lines themselves
    u128 seg_beta[QCF_FILTER_MAX_SEG] = {};
    u64 seg_ka[QCF_FILTER_MAX_SEG] = {};
    u64 seg_window[QCF_FILTER_MAX_SEG] = {};  // W_j: 0 <= cofactor - reference <= W_j for every guess of the segment
    // per-chain offset correction (FilterParams::rho)
    int mode = 0;                             // 0 constant floors, 1 drift lines, 2 drift lines + offset correction
    bool rho = false;
    u32 pinv32 = 0;
    u64 sigma0 = 0;
    u64 seg_sigma[QCF_FILTER_MAX_SEG] = {};
    u32 seg_kappa[QCF_FILTER_MAX_SEG] = {};
    double cost = 0;
};

} // namespace

QcfOptions qcf_options_snapshot()
{
    std::lock_guard<std::mutex> lock(g_opt_mu);
    return g_opt;
}

struct qcf_ctx {
    QcfOptions opt; // snapshot taken at the entry of the current ABI call
    unsigned ring_next = 0; // launches of the current call that took a row-group counter (QcfHits::work)
    FilterPlan last_plan;   // the largest filter launch (most periods) of the last sweep call: qcf_last_plan
    u32 last_plan_period = 0;
    u64 queue_spills = 0;   // survivors the filter launches of the last call tested on the spot (queue full): qcf_last_queue_spills
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    char name[128] = "";
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t flag_stream = nullptr;
    cudaStream_t aux[QCF_LANES - 1] = {}; // lanes 1..: lane 0 is `stream`
    cudaEvent_t ev_fork = nullptr, ev_lane[QCF_LANES - 1] = {};
    bool gcd_mode = false;                // the current sweep call runs the common-factor test (QCF_MODE_GCD)
    unsigned next_lane = 0;               // round-robin cursor of the current sweep call
    bool lane_used[QCF_LANES] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    QcfTables tables[QCF_TABLE_MAX_LEVEL + 1];
    bool have_tables[QCF_TABLE_MAX_LEVEL + 1] = {};
    u32 extra_exact = 0;                  // 29 while a level-10 sweep runs: the exact kernels drop its multiples
    u64 count_correction = 0;             // level-10 guesses the exact chain kernel counted although they are multiples of 29
    QcfHits* d_hits = nullptr;
    QcfHits* h_hits = nullptr; // pinned
    u32* h_flag = nullptr;     // pinned staging for the found flag
    u32* d_sink = nullptr;
    u32* d_dbg = nullptr; // QCF_DEBUG_STATS: survivor statistics of the filter kernel
    // peers connected through CUDA IPC: their found-flag words, mapped into this process
    int n_peers = 0;
    void* peer_base[64] = {};   // mapped QcfHits of each peer (for cudaIpcCloseMemHandle)
    u32** d_peer_flags = nullptr; // device array of &peer->stop
    int self_rank = 0;
    int peer_index_of_rank[64];   // rank -> index into peer_base (-1: self or not connected)
    unsigned long long* d_disp_out = nullptr; // result word of the dispenser kernel
    unsigned long long* h_disp_out = nullptr; // pinned
};

namespace {

struct RunStats {
    u64 launches = 0;
    double seconds = 0;
    u64 counted = 0;
    u32 kind = 0;
    int plan_degree = 0;
    u32 plan_tprime = 0, plan_fbits = 0;
};

// ---- lanes: the launches of one run_values call go round-robin to QCF_LANES streams ----------------------------
inline cudaStream_t lane_stream(qcf_ctx* ctx, unsigned lane) { return lane ? ctx->aux[lane - 1] : ctx->stream; }

// Everything enqueued on the main stream so far happens before anything the lanes run.
int lanes_fork(qcf_ctx* ctx)
{
    QCF_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (unsigned l = 1; l < QCF_LANES; ++l) {
        QCF_CUDA(cudaStreamWaitEvent(ctx->aux[l - 1], ctx->ev_fork, 0));
        ctx->lane_used[l] = false;
    }
    ctx->next_lane = 0;
    return QCF_OK;
}

inline unsigned lanes_next(qcf_ctx* ctx)
{
    const unsigned l = ctx->next_lane++ % QCF_LANES;
    ctx->lane_used[l] = true;
    return l;
}

// The main stream waits for every lane that was used.
int lanes_join(qcf_ctx* ctx)
{
    for (unsigned l = 1; l < QCF_LANES; ++l) {
        if (ctx->lane_used[l]) {
            QCF_CUDA(cudaEventRecord(ctx->ev_lane[l - 1], ctx->aux[l - 1]));
            QCF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane[l - 1], 0));
        }
    }
    return QCF_OK;
}

// Row-group counter of the next chain launch of this call: the ring is zeroed once per call (run_values); only a call with
// more than QCF_WORK_RING chain launches re-zeroes an entry, on the stream of its previous user.
int next_work_counter(qcf_ctx* ctx, cudaStream_t lane, unsigned long long** out)
{
    const unsigned i = ctx->ring_next++;
    *out = &ctx->d_hits->work[i % QCF_WORK_RING];
    if (i >= QCF_WORK_RING) {
        QCF_CUDA(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), lane));
    }
    return QCF_OK;
}

int ensure_tables(qcf_ctx* ctx, int level)
{
    if ((level < QCF_MIN_LEVEL) || (level > QCF_MAX_LEVEL)) {
        return fail(QCF_EINVAL, "level %d outside [5,10]", level);
    }
    level = table_level(level);
    if (!ctx->have_tables[level]) {
        QCF_CUDA(qcf_launch_build_tables(level, &ctx->tables[level], ctx->stream));
        QCF_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->have_tables[level] = true;
    }
    return QCF_OK;
}

// ---- filter-kernel planning ---------------------------------------------------------------------
// Cost of a plan in issue cycles per guess of a warp lane (calibrated on the B200 with tools/plan_cost.py, see
// profiles/): the trip body (one fused add+min per guess on the ALU pipe, D-1 multiply-adds on the FMA pipe, the
// trip's bookkeeping), the padding of the last trip, chain setup and segment re-basing amortised over the chain, and
// the replay of trips that contain a survivor (one lane of the warp verifies, the others wait).
struct FilterCostModel {
    // Fitted on a B200 to forced plans at four depths below sqrt(N) (tools/plan_cost.py).  Per segment: the re-basing, the slope
    // update of the drift lines (31 multiply-adds on the trip offsets) and the offset correction.
    // Refitted at the end of round 2 (gpurun_out/r2w_plan_cost.jsonl -> profiles/r02b_plan_cost.jsonl: 302 forced shapes at four
    // depths on the kernels with several trips per flag test and class-restricted replays; least squares over the shapes with >= 14
    // filter bits, rms 4.7 % at degree 2 packed).  What the earlier constants had wrong was the fixed cost of the packed degree-2
    // kernel: 306 per chain (two sets of packed offsets and moves) and 217 per segment (the base is rebuilt, the rest of a
    // group takes the slow path) -- the model said 200 and 100, and the planner bought filter bits with segments that cost more
    // than the survivors they saved (397 steps in 4 segments: 5.44 cycles per guess measured, 3.73 modelled).
    double body[4] = { 0, 1.29, 2.67, 4.69 };       // per guess at degree 1..3 (degree 1: packed 16-bit trips, four trips per flag test)
    double body2_packed = 1.99;                      // degree 2 in the packed form (sub-trips of 16 in groups of 8, two trips per flag test)
    double setup[4] = { 0, 173.0, 224.0, 210.0 };   // per chain
    double per_segment[4] = { 0, 93.0, 92.0, 85.0 }; // per segment of a chain (+ drift_segment + rho_segment, below)
    double setup2_packed = 306.0, per_segment2_packed = 177.0;
    double survivor = 13300.0;                       // x survivor rate (threshold / 2^32) per guess; degree 2 plain: x survivor2_plain; degree 3 (16-guess trips): x survivor3
    double survivor2_plain = 2.1, survivor3 = 2.4;
    double uncovered = 5.0;                          // what a plan leaves to the next pass (shorter chains) costs this much more per guess
    double drift_segment = 34.0;                     // the slope update of a segment (31 multiply-adds on the trip offsets)
    double rho_setup = 0.0, rho_segment = 6.0;       // per-chain offset correction: at chain setup (inside setup[]), per segment
    double replay16 = 0.03;                          // packed trips: x ((threshold >> 16) + 2 [+ 2 at degree 2: the groups' carries widen the high halves' threshold by G/2 = 4 (qcf_filter_math.cuh)]): the warp's replay of a flagged run of trips, per guess
};

FilterCostModel cost_model(const QcfOptions& o)
{
    FilterCostModel c;
    // qcf_set_option("cost_model", "body2,body3,setup2,setup3,segment2,segment3,survivor,uncovered"): 0 keeps the fitted value
    double* slot[8] = { &c.body[2], &c.body[3], &c.setup[2], &c.setup[3], &c.per_segment[2], &c.per_segment[3], &c.survivor, &c.uncovered };
    for (int i = 0; i < 8; ++i) {
        if (o.cost[i] > 0.0) {
            *slot[i] = o.cost[i];
        }
    }
    return c;
}

// Segment j of J gets trips / J trips, the top trips % J segments one more (the cofactor window of a segment grows with
// its length and with 1 / guess^2..3, so the longer ones go where the guesses are largest).
inline u32 seg_trips_of(u32 trips, u32 J, u32 j) { return trips / J + ((j >= J - trips % J) ? 1u : 0u); }
// (Segments of whole groups of the packed degree-2 kernel -- 4 trips, the odd trips in the last segment, so that only a chain's
// last segment ends in the kernel's slow path for runs shorter than a group -- were measured and not kept: slice 7 of 8 in 167.4 ms
// against 166.6, gpurun_out/r2z_ab.txt.)

// The reference line and the window of one segment: steps [k0, k1) of every chain of a launch whose chains start in
// (base, base + S) and step by S.  Step k of any chain is a guess g in (u_k, u_{k+1}), u_i = base + i*S, so its cofactor
// C = N/g (if g | N) lies strictly between f(k+1) and f(k), f(i) = N/u_i: convex and decreasing in i.  With F(i) = floor(f(i)):
//   * without drift the line is the constant F(k1) and the window F(k0) - F(k1): the cofactor's whole movement;
//   * with drift the line has the integer slope lambda <= f(ka+1) - f(ka+2) (the decrease per step at the middle ka,
//     rounded down: lambda = F(ka+1) - F(ka+2) - 1), so left of ka it stays below f(.+1) by convexity, and right of ka it is
//     lowered by 2 per step (the true decrease per step there is below lambda + 2):  l(k) = F(ka+1) - 2*(k1-1-ka) - lambda*(k-ka)
//     <= f(k+1) < C  for every k of the segment.  f(k) - l(k) is convex, so the window is taken at the two ends:
//     W = max(F(k0) + 1 - l(k0), F(k1-1) + 1 - l(k1-1)): the cofactor's CURVATURE over the segment + one step + ~2 per step.
// Exact integer arithmetic; returns false if the numbers leave the supported range.
struct SegLine {
    u64 lambda = 0, ka = 0;
    u128 beta = 0; // l(ka)
    u128 window = 0;
    u128 sigma = 0; // offset-correction mode: an upper bound of one step's cofactor drift at the segment's first step
};
inline u128 line_at(const SegLine& L, u64 k) { return (k >= L.ka) ? (L.beta - (u128)L.lambda * (k - L.ka)) : (L.beta + (u128)L.lambda * (L.ka - k)); }

// mode 0: constant floor; 1: drift line under f(k+1) (covers every chain offset: the window holds one whole step's drift);
// 2: drift line under f(k) + per-chain offset correction (qcf_filter_math.cuh): a chain at offset rho in (0, S) lowers its
//    reference by ~rho/S * sigma, sigma >= f(k0-1) - f(k0) >= S*N/u_k0^2 >= one step's drift anywhere in the segment.  With
//    C = N/(u_k + rho):  f(k) - rho/S * (f(k0-1) - f(k0)) <= C <= f(k) - rho/S * (f(ke) - f(ke+1))   (tangent below, chord above),
//    so the window holds the line's curvature part plus (sigma - (f(ke) - f(ke+1))): the CHANGE of a step's drift over the segment.
//    The caller adds the roundings of the correction arithmetic (plan_filter).
bool seg_line(u128 N, u128 base, u128 S, u64 k0, u64 k1, int mode, SegLine* out)
{
    if (!base || (k1 <= k0)) {
        return false;
    }
    const auto F = [&](u64 i) { return N / (base + (u128)i * S); };
    out->sigma = 0;
    if (mode == 0) {
        out->lambda = 0;
        out->ka = k0;
        out->beta = F(k1);
        out->window = F(k0) - out->beta;
        return true;
    }
    const u64 sh = (mode == 2) ? 0U : 1U; // the line runs under f(k + sh)
    const u64 ka = k0 + (k1 - k0) / 2U, ke = k1 - 1U;
    const u128 fa1 = F(ka + sh), fa2 = F(ka + sh + 1U);
    const u128 dec = fa1 - fa2; // within 1 of the real decrease per step at ka
    const u128 margin = (u128)2U * (ke - ka);
    if ((dec >> 62) || (fa1 <= margin + 2U)) {
        return false;
    }
    out->lambda = (dec > 0U) ? (u64)(dec - 1U) : 0U;
    out->ka = ka;
    out->beta = fa1 - margin;
    const u128 l0 = line_at(*out, k0), le = line_at(*out, ke);
    const u128 w0 = F(k0) + 1U - l0, we = F(ke) + 1U - le; // both lines lie below: no wrap
    out->window = std::max(w0, we);
    if (mode == 2) {
        if ((base <= S) && (k0 == 0U)) {
            return false; // no step below the first one to bound the drift with
        }
        const u128 below = (k0 == 0U) ? (N / (base - S)) : F(k0 - 1U);
        out->sigma = below - F(k0) + 1U;                 // >= f(k0-1) - f(k0)
        const u128 dlast = F(ke) - F(ke + 1U);           // <= f(ke) - f(ke+1) + 1
        const u128 slo = (dlast > 0U) ? (dlast - 1U) : 0U;
        if (out->sigma >> 62) {
            return false;
        }
        out->window += (out->sigma > slo) ? (out->sigma - slo) : 0U;
    }
    return true;
}

// Chooses the chain stride 2^tprime (in periods), the chain length K, the polynomial degree, the quotient alignment
// and the segmentation for whole periods starting at the first one inside [v_lo, v_hi].  The plan covers periods
// [m_lo, m_lo + (K << tprime)); the caller plans again for what is left above it.  Returns false when the filter does
// not apply (interval too short, cofactor range too wide for a 64-bit quotient, filter too weak).
// qcf_set_option("plan_force", "D,tprime,segments[,pad]") (0 = free) restricts the search (calibration and tests);
// "plan_drift" 0 plans with constant cofactor floors per segment instead of the drift lines.
bool plan_filter(const QcfOptions& opt, u128 N, const QcfTables& t, u128 v_lo, u128 v_hi, bool forced, FilterPlan* pl)
{
    const u128 P = t.period;
    const u128 m_lo = (v_lo + P - 1U) / P, m_max = (v_hi + 1U) / P;
    if ((m_max <= m_lo) || !m_lo) {
        return false;
    }
    const u128 M = m_max - m_lo;
    if (M >> 62) {
        return false;
    }
    int force_d = opt.force_d, force_tp = opt.force_tp, force_j = opt.force_j, force_pad = opt.force_pad; // force_pad: 1 = padded last trip, 2 = whole trips only
    if (force_tp && ((M >> force_tp) < 16U)) {
        force_d = force_tp = force_j = force_pad = 0; // the leftover of a forced launch is planned freely
    }
    const int want_mode = (opt.force_drift < 0) ? 2 : opt.force_drift; // 0 constant floors, 1 drift lines, 2 drift lines + per-chain offset correction
    const FilterCostModel cm = cost_model(opt);
    const u128 base = m_lo * P; // guesses lie strictly above
    const double Nd = std::ldexp((double)(u64)(N >> 64), 64) + (double)(u64)N, Pd = (double)t.period;
    const double v0d = std::ldexp((double)(u64)(base >> 64), 64) + (double)(u64)base;
    static const u32 seg_choices[6] = { 32, 16, 8, 4, 2, 1 };
    const double min_fail = forced ? 67108864.0 : 1048576.0; // bound + 1 >= 2^(32 - 6) resp. 2^(32 - 12): too few filter bits
    bool have = false;
    for (u32 tp = 0; tp <= 40; ++tp) {
        if ((M >> tp) < 16U) {
            break;
        }
        if ((P << tp) >> 62) {
            break;
        }
        if (force_tp && ((int)tp != force_tp)) {
            continue;
        }
        const int tt = (int)tp + 1; // 2-adic valuation of the stride (the period is even, once)
        const double two_tp = std::ldexp(1.0, (int)tp);
        const u128 S = P << tp;
        const double Sd = Pd * two_tp;
        // the offset correction needs a step below the first one (to bound the drift) and a 32-bit normalised offset
        const int mode = ((want_mode == 2) && ((base <= S) || (tp > 31U))) ? 1 : want_mode;
        const bool drift = mode != 0;
        const double seg_extra = drift ? (cm.drift_segment + ((mode == 2) ? cm.rho_segment : 0.0)) : 0.0; // slope (and offset) update of a segment, on top of the re-basing
        const double setup_extra = (mode == 2) ? cm.rho_setup : 0.0;
        u128 k128 = M >> tp;
        if (k128 >> 24) {
            k128 = (u128)1 << 24; // keeps the carry slack small; the rest is planned again
        }
        for (int variant = 0; variant < 6; ++variant) {
            // the whole span with the last trip padded past K, or whole trips only (what is left is planned again)
            const int D = variant / 2 + 1;
            if (force_d && (D != force_d)) {
                continue;
            }
            if (have && (std::min(cm.body[D], (D == 2) ? cm.body2_packed : cm.body[D]) >= pl->cost)) {
                continue; // the trip body alone costs more than the best plan (setup and survivors only add to it)
            }
            const u32 U = QCF_FILTER_UNROLL_FOR(D);
            // packed 16-bit trips: always at degree 1; at degree 2 when a sub-trip's move of the offsets, 16 * d2, is a multiple of
            // 2^16 (d2 is a multiple of 2^(33 - t), t = the stride's 2-adic valuation)
            const bool packed = (D == 1) || ((D == 2) && (tt <= 21) && (opt.force_packed != 0)) ;
            const double body_d = ((D == 2) && packed) ? cm.body2_packed : cm.body[D];
            const double setup_d = ((D == 2) && packed) ? cm.setup2_packed : cm.setup[D];
            const double seg_d = ((D == 2) && packed) ? cm.per_segment2_packed : cm.per_segment[D];
            const double surv_cost = cm.survivor * ((D >= 3) ? cm.survivor3 : (((D == 2) && !packed) ? cm.survivor2_plain : 1.0));
            u32 k = (u32)k128;
            if (force_pad && (force_pad != (variant & 1) + 1)) {
                continue;
            }
            if (variant & 1) {
                if (!force_pad && ((k % U == 0u) || (k < U))) {
                    continue;
                }
                k = k / U * U;
                if (!k) {
                    continue;
                }
            }
            const u32 trips = (k + U - 1U) / U, kpad = trips * U;
            const int w = std::min(64, (D + 1) * tt);
            const int a = 64 - w;
            if (a + 2 * tt < 32) {
                continue; // second and higher differences must live in the high limb only
            }
            const double kd = (double)k, covered = (double)(u64)((u128)k << tp) / (double)(u64)M;
            const double body_term = body_d * (double)kpad / kd;
            const double scale_a = std::ldexp(1.0, a - 32);
            if (have) {
                const double own_min = body_term + (setup_d + setup_extra + seg_d * 1U) / kd; // one segment: the least setup
                if (own_min * covered + (own_min + cm.uncovered) * (1.0 - covered) >= pl->cost) {
                    continue;
                }
            }
            u32 last_j = 0;
            for (u32 jc : seg_choices) {
                if (force_j && ((int)jc != force_j)) {
                    continue;
                }
                const u32 J = std::min(jc, trips);
                if (J == last_j) {
                    continue;
                }
                last_j = J;
                // what the tracked limb may run ahead of the true one: a carry out of the untracked low limb per step and per
                // re-basing, and (offset correction) the roundings of the per-segment correction: it may under-subtract by
                // rho_loss per segment (slack) and over-subtract by 2 (added to the thresholds)
                const u32 rho_loss = (mode == 2) ? ((a > 32) ? ((1U << (a - 32)) + 1U) : 2U) : 0U;
                const u32 slack = kpad - 1U + (J - 1U) * (1U + rho_loss);
                const u32 over = (mode == 2) ? 2U * (J - 1U) : 0U;
                const u32 t0 = seg_trips_of(trips, J, 0);
                const double own_fixed = body_term + (setup_d + setup_extra + (seg_d + seg_extra) * J - seg_extra) / kd;
                if (have && (own_fixed * covered + (own_fixed + cm.uncovered) * (1.0 - covered) >= pl->cost)) {
                    continue; // the survivor term can only add to it
                }
                // the widest window: the first segment (smallest guesses) or the first of the longer ones.  First in floating
                // point; the exact 128-bit divisions are made only for candidates that would beat the best plan so far.
                const u32 j1 = (trips % J) ? (J - trips % J) : 0U;
                {
                    const auto est = [&](double u0, double steps) {
                        const double span = steps * Sd;
                        if (!drift) {
                            return Nd * span / (u0 * (u0 + span));
                        }
                        const double curv = Nd * span * span / (4.0 * u0 * u0 * u0) * 0.9;
                        const double step = Nd * Sd / (u0 * (u0 + Sd)) * 0.9;
                        return curv + steps + ((mode == 2) ? step * std::min(1.0, 2.0 * span / (u0 + span)) * 0.9 : step);
                    };
                    double w_est = est(v0d, (double)t0 * U);
                    if (j1) {
                        w_est = std::max(w_est, est(v0d + ((double)j1 * t0 * U) * Sd, (double)(t0 + 1U) * U));
                    }
                    const double w_low = w_est * 0.99 - 2.0; // certainly not above the exact window
                    if ((w_low > 0.0) && (w_low * scale_a - 1.0 + slack + over >= min_fail)) {
                        continue;
                    }
                    if (have) {
                        const double own_est = own_fixed + surv_cost * (std::max(w_low, 0.0) * scale_a + slack + over) / 4294967296.0;
                        if (own_est * covered + (own_est + cm.uncovered) * (1.0 - covered) >= pl->cost) {
                            continue;
                        }
                    }
                }
                SegLine L;
                if (!seg_line(N, base, S, 0, (u64)t0 * U, mode, &L)) {
                    continue;
                }
                u128 W = L.window;
                const u128 rho_round = (mode == 2) ? (((L.sigma * ((P >> tp) + 2U)) >> 32) + 2U + J + 1U) : 0U; // see the final pass
                if (j1) {
                    const u64 ka = (u64)j1 * t0 * U;
                    if (!seg_line(N, base, S, ka, ka + (u64)(t0 + 1U) * U, mode, &L)) {
                        continue;
                    }
                    W = std::max(W, L.window);
                }
                W += rho_round;
                const int bitsW = bitlen(W);
                if ((bitsW > 62) || (w < bitsW)) {
                    continue;
                }
                const u128 B = (W << a) >> 32;
                const u128 bound = B + slack + over;
                if (bound >= 0xFFFFFFFEULL) {
                    continue;
                }
                const int fbits = 32 - bitlen(bound + 1U);
                if (fbits < (forced ? 6 : 12)) {
                    continue; // below 6 bits even one-trip epochs fill the survivor queue (128 lanes x 32 guesses x 2^-6 = 64 of 256)
                }
                const double own = own_fixed + surv_cost * (double)(u64)bound / 4294967296.0 + (packed ? cm.replay16 * ((double)(u64)(bound >> 16) + ((D == 2) ? 4.0 : 2.0)) : 0.0);
                const double cost = own * covered + (own + cm.uncovered) * (1.0 - covered);
                if (!have || (cost < pl->cost)) {
                    have = true;
                    pl->degree = D;
                    pl->packed = packed;
                    pl->tprime = tp;
                    pl->steps = k;
                    pl->trips = trips;
                    pl->n_seg = J;
                    pl->align = (u32)a;
                    pl->bound = (u32)bound;
                    pl->slack = slack;
                    pl->mode = mode;
                    pl->fbits = (u32)std::max(fbits, 0);
                    pl->cost = cost;
                }
            }
        }
    }
    if (!have) {
        return false;
    }
    // the lines, windows and re-basing constants of every segment of the chosen plan (the padded steps of the last trip
    // belong to the last segment: harmless)
    const u32 J = pl->n_seg, a = pl->align, U = QCF_FILTER_UNROLL_FOR(pl->degree);
    const int mode = pl->mode;
    const u128 S = P << pl->tprime;
    pl->drift = mode != 0;
    pl->rho = mode == 2;
    pl->pinv32 = (u32)(((u64)1 << 32) / t.period);
    pl->m_lo = m_lo;
    pl->m_end = m_lo + ((u128)pl->steps << pl->tprime);
    const u32 over = (mode == 2) ? 2U * (J - 1U) : 0U;
    SegLine prev;
    u128 sigma_prev = 0, rho_round = 0;
    u64 k0 = 0;
    u32 worst = 0;
    for (u32 j = 0; j < J; ++j) {
        pl->seg_trips[j] = seg_trips_of(pl->trips, J, j);
        const u64 k1 = k0 + (u64)pl->seg_trips[j] * U;
        SegLine L;
        if (!seg_line(N, base, S, k0, k1, mode, &L)) {
            return false; // cannot happen for the segments the search examined; the others are narrower
        }
        if (mode == 2) {
            // Offset correction, every rounding accounted for.  A chain at offset rho uses rho32 = floor(rho * 2^32 / S) - d,
            // 0 <= d < P / 2^tprime + 1, and lowers its reference by A_j: A_0 = floor(rho32 * sigma_0 / 2^32) + 1,
            // A_j = A_(j-1) - floor(rho32 * (sigma_(j-1) - sigma_j) / 2^32), hence
            //     rho/S * sigma_j - sigma_0 * (P / 2^tprime + 1) / 2^32  <  A_j  <=  rho/S * sigma_j + 1 + j.
            // The lines are lowered by the left deficit (rho_round holds it and the right excess), the sigmas kept
            // non-increasing (each is an upper bound, so the smaller of two is one too).
            if (j == 0) {
                rho_round = ((L.sigma * ((P >> pl->tprime) + 2U)) >> 32) + 2U;
                pl->sigma0 = (u64)L.sigma;
            } else if (L.sigma > sigma_prev) {
                L.sigma = sigma_prev;
            }
            if (L.beta <= rho_round + (u128)L.lambda * (k1 - L.ka) + 2U) {
                return false;
            }
            L.beta -= rho_round;
            L.window += rho_round + J + 1U;
        }
        if ((bitlen(L.window) > 64 - (int)a) || (bitlen(L.window) > 62)) {
            return false;
        }
        const u128 bj = ((L.window << a) >> 32) + pl->slack + over;
        if (bj >= 0xFFFFFFFEULL) {
            return false;
        }
        pl->seg_bound[j] = (u32)bj;
        worst = std::max(worst, pl->seg_bound[j]);
        pl->seg_lambda[j] = L.lambda;
        pl->seg_beta[j] = L.beta;
        pl->seg_ka[j] = L.ka;
        pl->seg_window[j] = (u64)L.window;
        pl->seg_sigma[j] = (u64)L.sigma;
        pl->seg_kappa[j] = 0;
        if (j == 0) {
            pl->c_lo = (u64)line_at(L, 0);   // the reference at step 0 (mod 2^64)
            pl->mu0 = (u64)L.lambda << a;    // mod 2^64
            pl->seg_shift[0] = 0;
            pl->seg_nu_lo[0] = pl->seg_nu_hi[0] = 0;
        } else {
            // entering segment j at step k0: R_j(k0) - R_{j-1}(k0) = (l_{j-1}(k0) - l_j(k0)) << a  (mod 2^64): the tracked
            // limb moves by its high limb (the carry from the low limb is in the slack); the first difference moves by
            // (lambda_j - lambda_{j-1}) << a, exactly, both limbs; the offset correction shrinks by rho32/2^32 * (sigma_(j-1) - sigma_j)
            const u64 delta = ((u64)line_at(prev, k0) - (u64)line_at(L, k0)) << a;
            pl->seg_shift[j] = (u32)(delta >> 32);
            const u64 nu = (L.lambda - prev.lambda) << a; // mod 2^64 (the slope shrinks up the chain: a negative number)
            pl->seg_nu_lo[j] = (u32)nu;
            pl->seg_nu_hi[j] = (u32)(nu >> 32);
            if (mode == 2) {
                pl->seg_kappa[j] = (u32)((((u128)(sigma_prev - L.sigma)) << a) >> 32); // < 2^32: sigma < 2^(64 - a)
            }
        }
        sigma_prev = L.sigma;
        prev = L;
        k0 = k1;
    }
    if (worst > pl->bound) { // a middle segment wider than the two the search examined (rounding of the lines): keep the books right
        pl->bound = worst;
        pl->fbits = (u32)std::max(0, 32 - bitlen((u128)worst + 1U));
        if (pl->fbits < (forced ? 6u : 12u)) {
            return false;
        }
    }
    return true;
}

int launch_exact_rows(qcf_ctx* ctx, u128 N, int ln, const QcfTables& t, u128 v_lo, u128 v_hi, bool count, u64* launches)
{
    if (v_hi < v_lo) {
        return QCF_OK;
    }
    const u128 m_lo = v_lo / t.period, m_hi = v_hi / t.period;
    const u128 top = (m_hi + 1U) * t.period; // largest value any thread forms, exclusive
    if (bitlen(top) > 96) {
        return fail(QCF_ERANGE, "guess values above 2^96 are not supported");
    }
    const u128 n_rows = (m_hi - m_lo + 1U) * t.n_b;
    if (n_rows >> 63) {
        return fail(QCF_ERANGE, "interval too long for one launch");
    }
    const int lg = std::max(1, (bitlen(top) + 31) / 32);
    const u128 c_max = N / ((v_lo < 1U) ? 1U : v_lo);
    const int qn = std::max(1, std::min((bitlen(c_max) + 31) / 32, ln));

    ExactParams prm;
    memset(&prm, 0, sizeof(prm));
    to_limbs(N, prm.n, 4);
    to_limbs(m_lo * t.period, prm.base, 3);
    to_limbs(v_lo, prm.v_lo, 3);
    to_limbs(v_hi, prm.v_hi, 3);
    prm.period = t.period;
    prm.n_a = t.n_a;
    prm.n_b = t.n_b;
    prm.m_span = (u64)(m_hi - m_lo);
    prm.extra_prime = ctx->extra_exact;
    prm.n_rows = (u64)n_rows;
    prm.d_a = t.d_a;
    prm.d_b = t.d_b;
    prm.hits = ctx->d_hits;
    const int grid = (int)std::min<u64>(prm.n_rows, (u64)ctx->sm_count * 4U);
    const unsigned lane = lanes_next(ctx);
    // a kernel exists for every (N limbs, guess limbs) the sweep of a semiprime needs; odd shapes (guesses wider
    // than half of N) run on the next wider N instantiation -- N's upper limbs are zero
    cudaError_t e = QCF_NO_KERNEL;
    for (int l = ln; (l <= QCF_MAX_N_LIMBS) && (e == QCF_NO_KERNEL); ++l) {
        if (ctx->gcd_mode) {
            e = qcf_launch_gcd_rows(prm, l, lg, count, grid, lane_stream(ctx, lane));
            continue;
        }
        for (int q = qn; (q <= l) && (e == QCF_NO_KERNEL); ++q) {
            e = qcf_launch_exact(prm, l, lg, q, count, grid, lane_stream(ctx, lane));
        }
    }
    if (e == QCF_NO_KERNEL) {
        return fail(QCF_ERANGE, "no exact kernel for N limbs %d, guess limbs %d, quotient words %d", ln, lg, qn);
    }
    QCF_CUDA(e);
    *launches += 1;
    return QCF_OK;
}

// One exact test per guess over [v_lo, v_hi]: whole periods as arithmetic-progression chains (qcf_exact_chain.cu: the
// index map is an add, the reciprocal is carried by one Newton step when the stride is a multiple of 2^16), ragged
// period ends and short intervals row by row (qcf_exact_kernel).
int launch_exact_values(qcf_ctx* ctx, u128 N, int ln, const QcfTables& t, u128 v_lo, u128 v_hi, bool count, u64* launches)
{
    if (v_hi < v_lo) {
        return QCF_OK;
    }
    const u128 P = t.period;
    const u128 m_lo = (v_lo + P - 1U) / P, m_max = (v_hi + 1U) / P;
    // Chains pay off only when there are enough of them to fill the GPU with stride >= 2^16 (a launch has
    // n_b << tprime rows, 4 rows per block): below 2^19 periods a chain launch is a few blocks walking long chains
    // (measured 60-85 us for 10^6 guesses), and the row kernel -- one guess per thread and row -- is far quicker.
    // (qcf_set_option("chain_min_log2") lowers the threshold: the chain form on intervals a CPU emulation can walk, tests/emu)
    const u128 chain_min = (u128)1 << ctx->opt.chain_min_log2;
    if ((m_max <= m_lo) || (m_max - m_lo < chain_min) || (bitlen(m_max * P) > 96) || ((m_max - m_lo) >> 62)) {
        return launch_exact_rows(ctx, N, ln, t, v_lo, v_hi, count, launches);
    }
    u128 cur = m_lo;
    for (int pass = 0; (pass < 4) && (m_max - cur >= chain_min); ++pass) {
        const u128 left = m_max - cur;
        const int bl = bitlen(left);
        // chains of up to 2^11 steps; common-factor mode: up to 2^13 (one gcd per chain, QCF_GCD_BATCH: the longer the chain, the
        // less the gcd costs per guess; there are still n_b * 2^15 * n_a >= 1.5e7 chains to fill the GPU with)
        const u32 tp = (u32)std::max(15, bl - (ctx->gcd_mode ? 13 : 11));
        const u128 k = left >> tp;
        const u128 g_min = (cur * P) ? (cur * P) : 1U;
        const int lg = std::max(1, (bitlen((cur + (k << tp)) * P) + 31) / 32);
        const int qn = std::max(1, std::min((bitlen(N / g_min) + 31) / 32, ln));
        FilterParams prm;
        memset(&prm, 0, sizeof(prm));
        to_limbs(N, prm.n, 4);
        to_limbs(cur * P, prm.base, 3);
        to_limbs(P << tp, prm.stride, 3);
        prm.n12p = (u64)(N >> 32) + (((u32)N != 0U) ? 1U : 0U);
        prm.period = t.period;
        prm.n_a = t.n_a;
        prm.n_b = t.n_b;
        prm.tprime = tp;
        prm.steps = (u32)k;
        prm.count = count ? 1u : 0u;
        if (ctx->extra_exact && !ctx->gcd_mode) {
            // the exact chain kernel runs at its register cap and knows nothing of prime 29: it counts the level-9
            // guesses of its periods and reports divisors that are multiples of 29 like any other; the host takes
            // both back (the multiples of 29 among the level-9 guesses of [lo, hi) are 29 * (level-9 guesses of [lo, hi) / 29))
            const u128 v0 = cur * P, v1 = (cur + (k << tp)) * P; // guesses lie strictly between
            const u32 q = ctx->extra_exact;
            ctx->count_correction += (u64)(count_coprime_upto((v1 - 1U) / q, QCF_TABLE_MAX_LEVEL) - count_coprime_upto(v0 / q, QCF_TABLE_MAX_LEVEL));
        } else if (ctx->extra_exact) {
            prm.extra[prm.n_extra++] = ctx->extra_exact;
        }
        prm.n_rows = (u64)t.n_b << tp;
        prm.nb_magic = (t.n_b > 1) ? (u64)(((u128)1 << 64) / t.n_b) + 1U : 0;
        prm.d_a = t.d_a;
        prm.d_b = t.d_b;
        prm.hits = ctx->d_hits;
        // the ring position decides the lane (entry i is always used on lane i % QCF_LANES)
        const unsigned lane = ctx->ring_next % QCF_LANES;
        ctx->lane_used[lane] = true;
        int rcw = next_work_counter(ctx, lane_stream(ctx, lane), &prm.work);
        if (rcw) {
            return rcw;
        }
        cudaError_t e = QCF_NO_KERNEL;
        for (int l = ln; (l <= QCF_MAX_N_LIMBS) && (e == QCF_NO_KERNEL); ++l) {
            if (ctx->gcd_mode) {
                // the Montgomery accumulator may stay unreduced (no subtraction per guess) when it cannot outgrow its l limbs:
                // every guess is below 2^gb and the top 32 lg - gb bits of N, as an l-limb number, are not all ones
                // (qcf_montmul_acc, qcf_gcd.cu)
                const int gb = bitlen((cur + (k << tp)) * P), sh = 32 * lg - gb;
                const bool lazy = (sh >= 1) && (sh < 32 * l) && ((N >> (32 * l - sh)) != (((u128)1 << sh) - 1U));
                e = qcf_launch_gcd_chain(prm, l, lg, lazy, ctx->sm_count, lane_stream(ctx, lane));
                continue;
            }
            for (int q = qn; (q <= l) && (e == QCF_NO_KERNEL); ++q) {
                e = qcf_launch_exact_chain(prm, l, lg, q, ctx->sm_count, ctx->opt.xchain_two_stage != 0, lane_stream(ctx, lane));
            }
        }
        if (e == QCF_NO_KERNEL) {
            break; // no chain instantiation for this shape: the row kernel takes the rest
        }
        QCF_CUDA(e);
        *launches += 1;
        cur += k << tp;
    }
    int rc = QCF_OK;
    if (v_lo < m_lo * P) {
        rc = launch_exact_rows(ctx, N, ln, t, v_lo, m_lo * P - 1U, count, launches);
    }
    if (!rc && (cur * P <= v_hi)) {
        rc = launch_exact_rows(ctx, N, ln, t, cur * P, v_hi, count, launches);
    }
    return rc;
}

// chain-draw random mode: which row groups of the plan a launch walks (drawn by Philox on the device)
struct RandomSelection {
    u64 seed = 0, block = 0;
    u32 first = 0, count = 0;
};

int launch_filter_values(qcf_ctx* ctx, u128 N, int ln, const QcfTables& t, const FilterPlan& pl, bool count, int base_level, int level,
    const RandomSelection* rs = nullptr)
{
    const u128 top = pl.m_end * t.period;
    if (bitlen(top) > 96) {
        return fail(QCF_ERANGE, "guess values above 2^96 are not supported");
    }
    const int lg = std::max(1, (bitlen(top) + 31) / 32);
    FilterParams prm;
    memset(&prm, 0, sizeof(prm));
    to_limbs(N, prm.n, 4);
    to_limbs(pl.m_lo * t.period, prm.base, 3);
    prm.period = t.period;
    prm.n_a = t.n_a;
    prm.n_b = t.n_b;
    prm.tprime = pl.tprime;
    prm.align = pl.align;
    prm.bound = pl.bound;
    prm.slack = pl.slack;
    prm.steps = pl.steps;
    prm.n_seg = pl.n_seg;
    for (u32 j = 0; j < pl.n_seg; ++j) {
        prm.seg_trips[j] = pl.seg_trips[j];
        prm.seg_shift[j] = pl.seg_shift[j];
        prm.seg_bound[j] = pl.seg_bound[j];
        prm.seg_nu_lo[j] = pl.seg_nu_lo[j];
        prm.seg_nu_hi[j] = pl.seg_nu_hi[j];
        prm.seg_kappa[j] = pl.seg_kappa[j];
    }
    prm.drift = pl.drift ? 1u : 0u;
    prm.mu0 = pl.mu0;
    prm.rho = pl.rho ? 1u : 0u;
    prm.pinv32 = pl.pinv32;
    prm.sigma0 = pl.sigma0;
    prm.count = count ? 1u : 0u;
    {
        // an epoch (the trips between two drains of the survivor queue) is sized for ~48 survivors per block: the queue
        // holds 256, and survivors that find it full anyway are tested on the spot by their lane (qcf_filter_spill)
        u32 worst = 1;
        for (u32 j = 0; j < pl.n_seg; ++j) {
            worst = std::max(worst, pl.seg_bound[j]);
        }
        const double per_trip = 128.0 * QCF_FILTER_UNROLL_FOR(pl.degree) * ((double)worst / 4294967296.0);
        const double trips = 48.0 / per_trip;
        prm.epoch_trips = (trips < 1.0) ? 1u : ((trips > 65536.0) ? 65536u : (u32)trips);
        if (ctx->opt.epoch_trips) { // qcf_set_option("epoch_trips"): the overflow test forces over-long epochs
            prm.epoch_trips = ctx->opt.epoch_trips;
        }
    }
    for (int q = base_level; q < level; ++q) {
        prm.extra[prm.n_extra++] = PRIMES10[q];
    }
    if (rs) {
        prm.rand_on = 1u;
        prm.rand_key0 = (u32)rs->seed;
        prm.rand_key1 = (u32)(rs->seed >> 32);
        prm.rand_block_lo = (u32)rs->block;
        prm.rand_block_hi = (u32)(rs->block >> 32);
        prm.rand_first = rs->first;
        prm.rand_count = rs->count;
    }
    prm.n64 = (u64)N;
    prm.c_lo = pl.c_lo;
    prm.n_rows = (u64)t.n_b << pl.tprime;
    prm.nb_magic = (t.n_b > 1) ? (u64)(((u128)1 << 64) / t.n_b) + 1U : 0;
    prm.d_a = t.d_a;
    prm.d_b = t.d_b;
    prm.hits = ctx->d_hits;
    if (ctx->opt.debug_stats) {
        if (!ctx->d_dbg) {
            QCF_CUDA(cudaMalloc(&ctx->d_dbg, sizeof(u32) * 4));
        }
        QCF_CUDA(cudaDeviceSynchronize()); // statistics are per launch: no overlap while they are collected
        QCF_CUDA(cudaMemsetAsync(ctx->d_dbg, 0, sizeof(u32) * 4, ctx->stream));
        QCF_CUDA(cudaStreamSynchronize(ctx->stream));
        prm.dbg = ctx->d_dbg;
    }
    const int grid = ctx->sm_count; // the launcher multiplies by the resident blocks per SM
    prm.rows_per_group = 4U; // fixed in the kernel (single-row claims for small launches measured no gain and cost the big ones 1 %)
    const unsigned lane = ctx->ring_next % QCF_LANES; // the ring position decides the lane
    ctx->lane_used[lane] = true;
    cudaStream_t ls = lane_stream(ctx, lane);
    {
        int rcw = next_work_counter(ctx, ls, &prm.work);
        if (rcw) {
            return rcw;
        }
    }
    cudaError_t e = QCF_NO_KERNEL;
    for (int l = ln; (l <= QCF_MAX_N_LIMBS) && (e == QCF_NO_KERNEL); ++l) {
        e = qcf_launch_filter(prm, pl.degree, pl.packed, l, lg, grid, ls);
    }
    if (e == QCF_NO_KERNEL) {
        return fail(QCF_ERANGE, "no filter kernel for N limbs %d, guess limbs %d", ln, lg);
    }
    QCF_CUDA(e);
    if (prm.dbg) {
        u32 h[4] = { 0, 0, 0, 0 };
        QCF_CUDA(cudaDeviceSynchronize());
        QCF_CUDA(cudaMemcpyAsync(h, ctx->d_dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        QCF_CUDA(cudaStreamSynchronize(ctx->stream));
        const double expect = (double)prm.n_rows * prm.n_a * prm.steps * ((double)pl.bound / 4294967296.0);
        fprintf(stderr, "[qcf] filter stats: %u survivors tested exactly (expected <= %.0f at segment 0's bound), epoch %u trips, fullest queue %u of %d; packed trips replayed by a lane: %u of %.0f\n",
            h[0], expect, prm.epoch_trips, h[1], 256, h[2], (double)prm.n_rows * prm.n_a * pl.trips);
    }
    return QCF_OK;
}

// The cheapest filter plan for the whole periods inside [v_lo, v_hi] over the wheel levels the chains may run on (superset
// mode: a LOWER level than requested, the level's remaining primes being applied to survivors; cost per level-L guess = plan
// cost x density ratio).  Returns the chains' level, 0 when the filter does not apply.
int best_filter_plan(qcf_ctx* ctx, u128 N, int level, u128 v_lo, u128 v_hi, bool forced, FilterPlan* pl)
{
    const int tl = table_level(level);
    int base_level = 0;
    double best = 0;
    // candidates: 5, 6, 7 and the level itself (native chains of levels 8 and 9 are a handful of steps long unless N is huge)
    for (int cand = (tl >= 6 ? 5 : tl); cand <= tl; cand = ((cand < 7) || (cand >= tl)) ? cand + 1 : tl) {
        FilterPlan p2;
        if ((ensure_tables(ctx, cand) != QCF_OK) || !plan_filter(ctx->opt, N, ctx->tables[cand], v_lo, v_hi, forced, &p2)) {
            continue;
        }
        double ratio = 1.0;
        for (int q = cand; q < level; ++q) {
            ratio *= (double)PRIMES10[q] / (double)(PRIMES10[q] - 1);
        }
        if (!base_level || (p2.cost * ratio < best)) {
            base_level = cand;
            best = p2.cost * ratio;
            *pl = p2;
        }
    }
    return base_level;
}

// Sweeps every guess of the level in the inclusive VALUE interval [v_lo, v_hi]: the whole wheel periods by the
// filter kernel when it applies (and is allowed), everything else by the exact kernel.
// On return ctx->h_hits holds the hit list of this interval.
int run_values(qcf_ctx* ctx, u128 N, int ln, int level, u128 v_lo, u128 v_hi, u32 flags, RunStats* st)
{
    if (v_hi < v_lo) {
        ctx->h_hits->count = 0;
        return QCF_OK;
    }
    int rc = ensure_tables(ctx, level);
    if (rc) {
        return rc;
    }
    const int tl = table_level(level);
    const QcfTables& t = ctx->tables[tl];
    ctx->extra_exact = (level > tl) ? PRIMES10[tl] : 0u;
    ctx->count_correction = 0;
    const bool count = (flags & QCF_COUNT_ON_DEVICE) != 0;
    const u32 kind = flags & QCF_KERNEL_MASK;
    // Common-factor mode has no filter (the 2-adic quotient only recognises exact divisors): every guess goes through
    // the gcd kernels.  They work modulo the odd part of N (Montgomery), which loses nothing: every guess is odd.
    ctx->gcd_mode = (flags & QCF_MODE_GCD) != 0;
    if (ctx->gcd_mode) {
        while (!(N & 1U)) {
            N >>= 1;
        }
        ln = std::max(1, (bitlen(N) + 31) / 32);
    }
    const bool allow_filter = (kind != QCF_KERNEL_EXACT) && !ctx->gcd_mode;
    const bool forced = kind == QCF_KERNEL_FILTER;

    // clear count, tested, overflow and the ring of row-group counters (one memset for the call, not one per launch); keep the stop flag
    QCF_CUDA(cudaMemsetAsync(&ctx->d_hits->count, 0, sizeof(u32), ctx->stream));
    static_assert(offsetof(QcfHits, work) == offsetof(QcfHits, tested) + sizeof(u64) + 2 * sizeof(u32), "tested, overflow, pad, work[] are contiguous");
    QCF_CUDA(cudaMemsetAsync(&ctx->d_hits->tested, 0, sizeof(u64) + 2 * sizeof(u32) + sizeof(unsigned long long) * QCF_WORK_RING, ctx->stream));
    ctx->ring_next = 0;
    QCF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = lanes_fork(ctx);
    if (rc) {
        return rc;
    }
    st->kind = QCF_KERNEL_EXACT;
    // The interval is cut into octaves [hi/2+1, hi] from the top: inside an octave the cofactor range
    // c_hi - c_lo stays below the cofactor itself, which is what the filter's strength depends on.  Each
    // octave is planned bottom-up (a plan covers whole periods from the first one; what is left above it is
    // planned again).  Whatever no plan covers -- ragged period ends, octaves far below sqrt(N) -- goes to
    // the exact kernel; once an octave cannot be planned at all, everything below it is one exact launch.
    u128 seg_hi = v_hi;
    bool filter_alive = allow_filter;
    FilterPlan pl;
    for (; !(ctx->gcd_mode && (N == 1U));) { // a power of two shares no factor with an odd guess
        u128 seg_lo = v_lo;
        if (filter_alive && ((seg_hi >> 1) + 1U > v_lo)) {
            seg_lo = (seg_hi >> 1) + 1U;
        }
        u128 cur = seg_lo; // everything of this segment below `cur` has been launched
        bool planned = false;
        for (int pass = 0; filter_alive && (pass < 8) && (cur <= seg_hi); ++pass) {
            // Chains may run over a LOWER wheel level than requested (superset mode): the filter is only a necessary
            // condition, and survivors are checked against the level's remaining primes before the exact test, so the
            // guesses that can be reported are exactly the level's.  Cost per level-L guess = plan cost x density ratio.
            const int base_level = best_filter_plan(ctx, N, level, cur, seg_hi, forced, &pl);
            if (!base_level) {
                break;
            }
            const QcfTables& tb = ctx->tables[base_level];
            planned = true;
            const u128 first = pl.m_lo * tb.period, last = pl.m_end * tb.period; // multiples of 2310: never guesses
            if (cur < first) {
                rc = launch_exact_values(ctx, N, ln, t, cur, first - 1U, count, &st->launches); // ragged bottom of the segment
                if (rc) {
                    return rc;
                }
            }
            rc = launch_filter_values(ctx, N, ln, tb, pl, count, base_level, level);
            if (rc) {
                return rc;
            }
            st->launches += 1;
            if (ctx->opt.debug) {
                fprintf(stderr, "[qcf] filter plan: level %d (chains at level %d) periods %llu degree %d tprime %u steps %u segments %u align %u bound %u (filter bits %u) cost %.2f\n",
                    level, base_level, (unsigned long long)(u64)(pl.m_end - pl.m_lo), pl.degree, pl.tprime, pl.steps, pl.n_seg, pl.align, pl.bound, pl.fbits, pl.cost);
            }
            st->kind = QCF_KERNEL_FILTER;
            if (!ctx->last_plan_period || ((pl.m_end - pl.m_lo) * tb.period > (ctx->last_plan.m_end - ctx->last_plan.m_lo) * ctx->last_plan_period)) {
                ctx->last_plan = pl;
                ctx->last_plan_period = tb.period;
            }
            st->plan_degree = pl.degree;
            st->plan_tprime = pl.tprime;
            st->plan_fbits = pl.fbits;
            cur = last;
        }
        if (!planned && (seg_lo > v_lo)) {
            filter_alive = false; // no plan for this octave: the rest of the interval goes to the exact kernel
            continue;
        }
        if (cur <= seg_hi) {
            rc = launch_exact_values(ctx, N, ln, t, cur, seg_hi, count, &st->launches); // what the filter does not cover
            if (rc) {
                return rc;
            }
        }
        if (seg_lo <= v_lo) {
            break;
        }
        seg_hi = seg_lo - 1U;
    }
    rc = lanes_join(ctx);
    if (rc) {
        return rc;
    }
    QCF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    QCF_CUDA(cudaMemcpyAsync(ctx->h_hits, ctx->d_hits, QCF_HITS_D2H_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    QCF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    st->seconds += ms * 1e-3;
    ctx->queue_spills += ctx->h_hits->overflow; // survivors tested by their own lane because the queue was full (qcf_filter_spill)
    if (ctx->extra_exact && ctx->h_hits->count) {
        // level 10: drop reported divisors that are multiples of 29 (see launch_exact_values)
        QcfHits* h = ctx->h_hits;
        const u32 nh = std::min<u32>(h->count, QCF_MAX_HITS);
        u32 kept = 0;
        for (u32 i = 0; i < nh; ++i) {
            if (from_limbs(h->g[i], 4) % ctx->extra_exact) {
                if (kept != i) {
                    memcpy(h->g[kept], h->g[i], sizeof(h->g[i]));
                }
                ++kept;
            }
        }
        if (h->count <= QCF_MAX_HITS) {
            h->count = kept;
        }
    }
    st->counted += ctx->h_hits->tested - (count ? ctx->count_correction : 0U);
    return QCF_OK;
}

// Sweeps the inclusive index interval; fills `out` (accumulating counters) and picks the reported divisor.
int run_indices(qcf_ctx* ctx, u128 N, int ln, int level, u128 p_lo, u128 p_hi, u32 flags, qcf_result* out)
{
    if (p_hi < p_lo) {
        return QCF_OK;
    }
    const u128 v_lo = h_forward(p_lo), v_hi = h_forward(p_hi);
    RunStats st;
    int rc = run_values(ctx, N, ln, level, v_lo, v_hi, flags, &st);
    if (rc) {
        return rc;
    }
    out->kernel_launches += st.launches;
    out->device_seconds += st.seconds;
    out->kernel_kind = st.kind;
    if (ctx->h_hits->stop) {
        out->aborted = 1;
        return QCF_OK;
    }
    out->guesses_counted += st.counted;
    out->guesses_tested += (u64)(count_coprime_upto(v_hi, level) - count_coprime_upto(v_lo - 1U, level));
    out->n_hits = ctx->h_hits->count;
    u128 lo = p_lo, hi = p_hi;
    // More hits than the record holds (a number with many divisors; in common-factor mode any N with a smallish prime
    // factor): the record is then an arbitrary subset, so the interval is narrowed -- to its highest batch, then by
    // halving towards the lower indices, the reference's visiting order -- until the record is complete.
    // Invariant: the first hit in visiting order lies in [lo, hi], and the record describes [lo, hi].
    for (int round = 0; (ctx->h_hits->count > QCF_MAX_HITS) && (round < 400); ++round) {
        const u128 top_lo = batch_p_lo((u64)batch_of_index(hi));
        const bool by_batch = lo < top_lo;
        const u128 cut = by_batch ? (top_lo - 1U) : (lo + (hi - lo) / 2U); // [lo, cut] | [cut + 1, hi]
        // visited first: the highest batch [cut + 1, hi] when cutting by batch, the lower half [lo, cut] inside a batch
        const u128 a_lo = by_batch ? (cut + 1U) : lo, a_hi = by_batch ? hi : cut;
        rc = run_values(ctx, N, ln, level, h_forward(a_lo), h_forward(a_hi), flags, &st);
        if (!rc && !ctx->h_hits->stop) {
            if (ctx->h_hits->count) {
                lo = a_lo;
                hi = a_hi;
                continue;
            }
            // nothing there: the first hit is in the other part
            if (by_batch) {
                hi = cut;
            } else {
                lo = cut + 1U;
            }
            rc = run_values(ctx, N, ln, level, h_forward(lo), h_forward(hi), flags, &st);
        }
        if (rc) {
            return rc;
        }
        if (ctx->h_hits->stop) {
            out->aborted = 1;
            return QCF_OK;
        }
    }
    const u32 nh = std::min<u32>(ctx->h_hits->count, QCF_MAX_HITS);
    if (nh) {
        // the reference's single-thread order: highest batch first, ascending inside a batch
        u128 best = 0, best_batch = 0;
        for (u32 i = 0; i < nh; ++i) {
            const u128 g = from_limbs(ctx->h_hits->g[i], 4);
            const u128 b = batch_of_index(h_backward(g) - 1U);
            if ((i == 0) || (b > best_batch) || ((b == best_batch) && (g < best))) {
                best = g;
                best_batch = b;
            }
        }
        out->found = 1;
        if (flags & QCF_MODE_GCD) {
            // what the reference prints first: gcd(base, toFactor) (include/qimcifa.hpp:374-376)
            u128 a = best, b = N;
            while (b) {
                const u128 r = a % b;
                a = b;
                b = r;
            }
            best = a;
        }
        to_limbs(best, out->factor_le, 4);
    }
    return QCF_OK;
}

void export_plan(const FilterPlan& pl, u32 period, qcf_filter_plan* out)
{
    memset(out, 0, sizeof(*out));
    out->period = period;
    out->found = 1;
    out->degree = pl.degree;
    out->tprime = pl.tprime;
    out->align = pl.align;
    out->steps = pl.steps;
    out->trips = pl.trips;
    out->unroll = QCF_FILTER_UNROLL_FOR(pl.degree);
    out->n_seg = pl.n_seg;
    out->slack = pl.slack;
    out->fbits = pl.fbits;
    for (u32 j = 0; j < pl.n_seg; ++j) {
        out->seg_trips[j] = pl.seg_trips[j];
        out->seg_shift[j] = pl.seg_shift[j];
        out->seg_bound[j] = pl.seg_bound[j];
        out->seg_nu_lo[j] = pl.seg_nu_lo[j];
        out->seg_nu_hi[j] = pl.seg_nu_hi[j];
        out->seg_lambda[j] = pl.seg_lambda[j];
        out->seg_ka[j] = pl.seg_ka[j];
        out->seg_window[j] = pl.seg_window[j];
        to_limbs(pl.seg_beta[j], out->seg_beta_le[j], 4);
        out->seg_sigma[j] = pl.seg_sigma[j];
        out->seg_kappa[j] = pl.seg_kappa[j];
    }
    out->drift = pl.drift ? 1u : 0u;
    out->mu0 = pl.mu0;
    out->rho = pl.rho ? 1u : 0u;
    out->pinv32 = pl.pinv32;
    out->sigma0 = pl.sigma0;
    to_limbs(pl.m_lo, out->m_lo_le, 4);
    to_limbs(pl.m_end, out->m_end_le, 4);
    out->c_lo = pl.c_lo;
    out->cost = pl.cost;
    out->packed = pl.packed ? 1u : 0u;
}

int check_n(const u32* n_le, int n_limbs, u128* N, int* ln)
{
    if (!n_le || (n_limbs < 1) || (n_limbs > QCF_MAX_N_LIMBS)) {
        return fail(QCF_EINVAL, "n_limbs must be in [1,%d]", QCF_MAX_N_LIMBS);
    }
    *N = from_limbs(n_le, n_limbs);
    if (*N < 2) {
        return fail(QCF_EINVAL, "number to factor must be >= 2");
    }
    *ln = std::max(2, (bitlen(*N) + 31) / 32);
    return QCF_OK;
}

} // namespace

extern "C" {

const char* qcf_last_error(void) { return g_err; }

int qcf_create(int device, qcf_ctx** out)
{
    if (!out) {
        return fail(QCF_EINVAL, "out is null");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if ((e != cudaSuccess) || (ndev == 0)) {
        return fail(QCF_ENODEV, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    }
    if ((device < 0) || (device >= ndev)) {
        return fail(QCF_EINVAL, "device %d out of range [0,%d)", device, ndev);
    }
    QCF_CUDA(cudaSetDevice(device));
    qcf_ctx* ctx = new qcf_ctx();
    ctx->device = device;
    const int rc = [ctx, device]() -> int {
        cudaDeviceProp prop;
        QCF_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        QCF_CUDA(cudaDeviceGetAttribute(&ctx->clock_khz, cudaDevAttrClockRate, device));
        snprintf(ctx->name, sizeof(ctx->name), "%s", prop.name);
        QCF_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
        QCF_CUDA(cudaStreamCreateWithFlags(&ctx->flag_stream, cudaStreamNonBlocking));
        QCF_CUDA(cudaEventCreate(&ctx->ev0));
        QCF_CUDA(cudaEventCreate(&ctx->ev1));
        QCF_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (unsigned l = 0; l + 1 < QCF_LANES; ++l) {
            QCF_CUDA(cudaStreamCreateWithFlags(&ctx->aux[l], cudaStreamNonBlocking));
            QCF_CUDA(cudaEventCreateWithFlags(&ctx->ev_lane[l], cudaEventDisableTiming));
        }
        QCF_CUDA(cudaMalloc(&ctx->d_hits, sizeof(QcfHits)));
        QCF_CUDA(cudaMemset(ctx->d_hits, 0, sizeof(QcfHits)));
        QCF_CUDA(cudaMallocHost(&ctx->h_hits, sizeof(QcfHits)));
        memset(ctx->h_hits, 0, sizeof(QcfHits));
        QCF_CUDA(cudaMallocHost(&ctx->h_flag, sizeof(u32)));
        QCF_CUDA(cudaMalloc(&ctx->d_sink, sizeof(u32)));
        return QCF_OK;
    }();
    if (rc) {
        qcf_destroy(ctx); // releases whatever was created before the failure (every handle starts out null)
        return rc;
    }
    *out = ctx;
    return QCF_OK;
}

int qcf_destroy(qcf_ctx* ctx)
{
    if (!ctx) {
        return QCF_OK;
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int l = 0; l <= QCF_TABLE_MAX_LEVEL; ++l) {
        if (ctx->have_tables[l]) {
            cudaFree(ctx->tables[l].d_a);
            cudaFree(ctx->tables[l].d_b);
        }
    }
    qcf_peer_disconnect(ctx);
    cudaFree(ctx->d_hits);
    cudaFree(ctx->d_sink);
    cudaFree(ctx->d_dbg);
    cudaFree(ctx->d_disp_out);
    cudaFreeHost(ctx->h_disp_out);
    cudaFreeHost(ctx->h_hits);
    cudaFreeHost(ctx->h_flag);
    if (ctx->ev0) {
        cudaEventDestroy(ctx->ev0);
    }
    if (ctx->ev1) {
        cudaEventDestroy(ctx->ev1);
    }
    if (ctx->ev_fork) {
        cudaEventDestroy(ctx->ev_fork);
    }
    for (unsigned l = 0; l + 1 < QCF_LANES; ++l) {
        if (ctx->aux[l]) {
            cudaStreamDestroy(ctx->aux[l]);
        }
        if (ctx->ev_lane[l]) {
            cudaEventDestroy(ctx->ev_lane[l]);
        }
    }
    if (ctx->own_stream && ctx->stream) {
        cudaStreamDestroy(ctx->stream);
    }
    if (ctx->flag_stream) {
        cudaStreamDestroy(ctx->flag_stream);
    }
    delete ctx;
    return QCF_OK;
}

int qcf_set_stream(qcf_ctx* ctx, void* cuda_stream)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    ctx->stream = (cudaStream_t)cuda_stream;
    return QCF_OK;
}

int qcf_device_info(qcf_ctx* ctx, int* sm_count, int* clock_khz, char* name, int name_len)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    if (sm_count) {
        *sm_count = ctx->sm_count;
    }
    if (clock_khz) {
        *clock_khz = ctx->clock_khz;
    }
    if (name && (name_len > 0)) {
        snprintf(name, (size_t)name_len, "%s", ctx->name);
    }
    return QCF_OK;
}

int qcf_build_tables(qcf_ctx* ctx, int level)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    return ensure_tables(ctx, level);
}

int qcf_get_tables(qcf_ctx* ctx, int level, uint32_t* period, uint32_t* n_a, uint32_t* n_b, uint32_t* table_a, uint32_t cap_a,
    uint32_t* table_b, uint32_t cap_b)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    if (level > QCF_TABLE_MAX_LEVEL) {
        return fail(QCF_EINVAL, "level 10 has no tables of its own (it runs on the level-9 tables)");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_tables(ctx, level);
    if (rc) {
        return rc;
    }
    const QcfTables& t = ctx->tables[level];
    if (period) {
        *period = t.period;
    }
    if (n_a) {
        *n_a = t.n_a;
    }
    if (n_b) {
        *n_b = t.n_b;
    }
    if (table_a) {
        if (cap_a < t.n_a) {
            return fail(QCF_EINVAL, "table_a capacity %u < %u", cap_a, t.n_a);
        }
        QCF_CUDA(cudaMemcpy(table_a, t.d_a, sizeof(u32) * t.n_a, cudaMemcpyDeviceToHost));
    }
    if (table_b) {
        if (cap_b < t.n_b) {
            return fail(QCF_EINVAL, "table_b capacity %u < %u", cap_b, t.n_b);
        }
        QCF_CUDA(cudaMemcpy(table_b, t.d_b, sizeof(u32) * t.n_b, cudaMemcpyDeviceToHost));
    }
    return QCF_OK;
}

int qcf_sweep(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t batch_lo, uint64_t batch_hi,
    uint64_t batches_per_launch, uint32_t flags, qcf_result* out)
{
    if (!ctx || !out) {
        return fail(QCF_EINVAL, "ctx/out is null");
    }
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    if (batch_hi < batch_lo) {
        return fail(QCF_EINVAL, "batch_hi < batch_lo");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    ctx->opt = qcf_options_snapshot();
    ctx->last_plan_period = 0;
    ctx->queue_spills = 0;
    memset(out, 0, sizeof(*out));
    out->next_batch_hi = batch_hi;
    const u64 step = batches_per_launch ? batches_per_launch : 1024U;
    u64 hi = batch_hi;
    while (hi > batch_lo) {
        const u64 lo = (hi - batch_lo > step) ? (hi - step) : batch_lo;
        rc = run_indices(ctx, N, ln, level, batch_p_lo(lo), batch_p_hi(hi - 1U), flags, out);
        if (rc) {
            return rc;
        }
        if (out->aborted) {
            break;
        }
        out->batches_done += hi - lo;
        hi = lo;
        out->next_batch_hi = hi;
        if (out->found) {
            if (ctx->n_peers) {
                qcf_peer_signal(ctx); // peers stop at their next poll (finish(), reference include/qimcifa.hpp:102-105)
            }
            break;
        }
    }
    return QCF_OK;
}

int qcf_sweep_indices(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, const uint32_t* p_lo_le,
    const uint32_t* p_hi_le, uint32_t flags, qcf_result* out)
{
    if (!ctx || !out || !p_lo_le || !p_hi_le) {
        return fail(QCF_EINVAL, "null argument");
    }
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    ctx->opt = qcf_options_snapshot();
    ctx->last_plan_period = 0;
    ctx->queue_spills = 0;
    memset(out, 0, sizeof(*out));
    return run_indices(ctx, N, ln, level, from_limbs(p_lo_le, 4), from_limbs(p_hi_le, 4), flags, out);
}

int qcf_last_hits(qcf_ctx* ctx, uint32_t* hits_le, uint32_t cap, uint32_t* n_hits)
{
    if (!ctx || !n_hits) {
        return fail(QCF_EINVAL, "null argument");
    }
    const u32 nh = std::min<u32>(ctx->h_hits->count, QCF_MAX_HITS);
    *n_hits = nh;
    for (u32 i = 0; (i < nh) && (i < cap) && hits_le; ++i) {
        memcpy(hits_le + 4 * i, ctx->h_hits->g[i], 16);
    }
    return QCF_OK;
}

int qcf_node_split(const uint32_t* n_le, int n_limbs, uint64_t node_count, uint32_t* sqrt_le, uint32_t* full_range_le,
    uint64_t* node_range, uint64_t* batch_count)
{
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    if (!node_count) {
        return fail(QCF_EINVAL, "node_count must be >= 1");
    }
    const u128 s = h_isqrt(N);
    const u128 fullRange = h_backward(s);
    const u128 nodeRange = (((fullRange + node_count - 1U) / node_count) + QCF_BIGGEST_WHEEL - 1U) / QCF_BIGGEST_WHEEL;
    const u128 bc = nodeRange * node_count;
    if (bc >> 64) {
        return fail(QCF_ERANGE, "batch count exceeds 64 bits");
    }
    if (sqrt_le) {
        to_limbs(s, sqrt_le, 4);
    }
    if (full_range_le) {
        to_limbs(fullRange, full_range_le, 4);
    }
    if (node_range) {
        *node_range = (u64)nodeRange;
    }
    if (batch_count) {
        *batch_count = (u64)bc;
    }
    return QCF_OK;
}

int qcf_count_guesses(int level, uint64_t batch_lo, uint64_t batch_hi, uint32_t* count_le)
{
    if ((level < QCF_MIN_LEVEL) || (level > QCF_MAX_LEVEL) || !count_le) {
        return fail(QCF_EINVAL, "bad level or null output");
    }
    u128 c = 0;
    if (batch_hi > batch_lo) {
        c = count_coprime_upto(h_forward(batch_p_hi(batch_hi - 1U)), level) - count_coprime_upto(h_forward(batch_p_lo(batch_lo)) - 1U, level);
    }
    to_limbs(c, count_le, 4);
    return QCF_OK;
}

int qcf_set_found(qcf_ctx* ctx, int value)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    *ctx->h_flag = value ? 1u : 0u;
    QCF_CUDA(cudaMemcpyAsync(&ctx->d_hits->stop, ctx->h_flag, sizeof(u32), cudaMemcpyHostToDevice, ctx->flag_stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->flag_stream));
    return QCF_OK;
}

int qcf_peer_export(qcf_ctx* ctx, uint8_t* handle)
{
    if (!ctx || !handle) {
        return fail(QCF_EINVAL, "null argument");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) <= QCF_PEER_HANDLE_BYTES, "handle size");
    QCF_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    QCF_CUDA(cudaIpcGetMemHandle(&h, ctx->d_hits));
    memset(handle, 0, QCF_PEER_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    return QCF_OK;
}

int qcf_peer_disconnect(qcf_ctx* ctx)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    cudaSetDevice(ctx->device);
    if (ctx->flag_stream) {
        cudaStreamSynchronize(ctx->flag_stream);
    }
    for (int i = 0; i < ctx->n_peers; ++i) {
        cudaIpcCloseMemHandle(ctx->peer_base[i]);
    }
    ctx->n_peers = 0;
    if (ctx->d_peer_flags) {
        cudaFree(ctx->d_peer_flags);
        ctx->d_peer_flags = nullptr;
    }
    return QCF_OK;
}

int qcf_peer_connect(qcf_ctx* ctx, const uint8_t* handles, int n_ranks, int self_rank)
{
    if (!ctx || !handles || (n_ranks < 1) || (n_ranks > 64) || (self_rank < 0) || (self_rank >= n_ranks)) {
        return fail(QCF_EINVAL, "bad peer arguments (at most 64 ranks)");
    }
    qcf_peer_disconnect(ctx);
    QCF_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < 64; ++r) {
        ctx->peer_index_of_rank[r] = -1;
    }
    u32* flags[64];
    int n = 0;
    for (int r = 0; r < n_ranks; ++r) {
        if (r == self_rank) {
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * QCF_PEER_HANDLE_BYTES, sizeof(h));
        void* base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            ctx->n_peers = n;
            qcf_peer_disconnect(ctx);
            return fail(QCF_ECUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
        }
        ctx->peer_base[n] = base;
        ctx->peer_index_of_rank[r] = n;
        flags[n] = &((QcfHits*)base)->stop;
        ++n;
    }
    ctx->n_peers = n;
    ctx->self_rank = self_rank;
    if (n) {
        QCF_CUDA(cudaMalloc(&ctx->d_peer_flags, sizeof(u32*) * n));
        QCF_CUDA(cudaMemcpy(ctx->d_peer_flags, flags, sizeof(u32*) * n, cudaMemcpyHostToDevice));
    }
    return QCF_OK;
}

int qcf_peer_signal(qcf_ctx* ctx)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    if (!ctx->n_peers) {
        return QCF_OK;
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    QCF_CUDA(qcf_launch_signal_peers(ctx->d_peer_flags, ctx->n_peers, ctx->flag_stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->flag_stream));
    return QCF_OK;
}

int qcf_dispenser_reset(qcf_ctx* ctx, uint64_t value)
{
    if (!ctx) {
        return fail(QCF_EINVAL, "ctx is null");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    QCF_CUDA(cudaMemcpyAsync(&ctx->d_hits->dispenser, &value, sizeof(value), cudaMemcpyHostToDevice, ctx->flag_stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->flag_stream));
    return QCF_OK;
}

int qcf_peer_dispense(qcf_ctx* ctx, int owner_rank, uint64_t take, uint64_t* start)
{
    if (!ctx || !start || (owner_rank < 0) || (owner_rank >= 64)) {
        return fail(QCF_EINVAL, "bad dispenser arguments");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* word = nullptr;
    if ((owner_rank == ctx->self_rank) || !ctx->n_peers) {
        word = &ctx->d_hits->dispenser;
    } else {
        const int idx = ctx->peer_index_of_rank[owner_rank];
        if ((idx < 0) || (idx >= ctx->n_peers)) {
            return fail(QCF_EINVAL, "rank %d is not a connected peer", owner_rank);
        }
        word = &((QcfHits*)ctx->peer_base[idx])->dispenser;
    }
    if (!ctx->d_disp_out) {
        QCF_CUDA(cudaMalloc(&ctx->d_disp_out, sizeof(unsigned long long)));
        QCF_CUDA(cudaMallocHost(&ctx->h_disp_out, sizeof(unsigned long long)));
    }
    QCF_CUDA(qcf_launch_dispense(word, take, ctx->d_disp_out, ctx->flag_stream));
    QCF_CUDA(cudaMemcpyAsync(ctx->h_disp_out, ctx->d_disp_out, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->flag_stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->flag_stream));
    *start = *ctx->h_disp_out;
    return QCF_OK;
}

int qcf_poll_found(qcf_ctx* ctx, int* value)
{
    if (!ctx || !value) {
        return fail(QCF_EINVAL, "null argument");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    u32 v = 0;
    QCF_CUDA(cudaMemcpyAsync(&v, &ctx->d_hits->stop, sizeof(u32), cudaMemcpyDeviceToHost, ctx->flag_stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->flag_stream));
    *value = (int)v;
    return QCF_OK;
}

int qcf_time_level(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t batches, uint32_t flags,
    double* seconds, uint64_t* guesses)
{
    if (!ctx || !seconds || !guesses) {
        return fail(QCF_EINVAL, "null argument");
    }
    uint64_t node_range = 0, batch_count = 0;
    int rc = qcf_node_split(n_le, n_limbs, 1, nullptr, nullptr, &node_range, &batch_count);
    if (rc) {
        return rc;
    }
    if (!batches) {
        batches = 1;
    }
    // the batches just below sqrt(N): the reference tuner's offset (src/qimcifa_tuner.cpp:70-71) aims there
    const u64 hi = batch_count;
    const u64 lo = (batch_count > batches) ? (batch_count - batches) : 0;
    u128 N;
    int ln;
    rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    ctx->opt = qcf_options_snapshot();
    ctx->last_plan_period = 0;
    ctx->queue_spills = 0;
    qcf_result res;
    memset(&res, 0, sizeof(res));
    // one launch over the whole span; hits are ignored (the tuner only times)
    rc = run_indices(ctx, N, ln, level, batch_p_lo(lo), batch_p_hi(hi - 1U), flags, &res);
    if (rc) {
        return rc;
    }
    if (res.aborted) {
        // the found flag is sticky (qcf_set_found(ctx, 0) clears it): a timing of a sweep that stopped early would be a cost of 0
        return fail(QCF_EINVAL, "the found flag is raised: the timed sweep was aborted (clear it with qcf_set_found(ctx, 0))");
    }
    *seconds = res.device_seconds;
    *guesses = res.guesses_tested;
    return QCF_OK;
}

static int run_random(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t budget, u32* d_dump, RunStats* st)
{
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    if (level > QCF_TABLE_MAX_LEVEL) {
        return fail(QCF_EINVAL, "random-guess mode draws from the residue tables: levels 5..9");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    rc = ensure_tables(ctx, level);
    if (rc) {
        return rc;
    }
    const QcfTables& t = ctx->tables[level];
    const u128 s = h_isqrt(N);
    const u128 M = (s + 1U) / t.period;
    if (M == 0) {
        return fail(QCF_ERANGE, "sqrt(N) is below one wheel period of level %d; use the sweep", level);
    }
    if (M >> 64) {
        return fail(QCF_ERANGE, "period count exceeds 64 bits");
    }
    const u128 top = M * t.period;
    const int lg = std::max(1, (bitlen(top) + 31) / 32);
    RandomParams prm;
    memset(&prm, 0, sizeof(prm));
    to_limbs(N, prm.n, 4);
    prm.period = t.period;
    prm.n_a = t.n_a;
    prm.n_b = t.n_b;
    prm.key0 = (u32)seed;
    prm.key1 = (u32)(seed >> 32);
    prm.periods = (u64)M;
    prm.counter_lo = counter_lo;
    prm.budget = budget;
    prm.d_a = t.d_a;
    prm.d_b = t.d_b;
    prm.hits = ctx->d_hits;
    prm.dump = d_dump;
    const u64 want = (budget + 255U) / 256U;
    const int grid = (int)std::max<u64>(1, std::min<u64>(want, (u64)ctx->sm_count * 8U));
    QCF_CUDA(cudaMemsetAsync(&ctx->d_hits->count, 0, sizeof(u32), ctx->stream));
    QCF_CUDA(cudaMemsetAsync(&ctx->d_hits->tested, 0, sizeof(u64), ctx->stream));
    QCF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    cudaError_t e = QCF_NO_KERNEL;
    for (int l = ln; (l <= QCF_MAX_N_LIMBS) && (e == QCF_NO_KERNEL); ++l) {
        e = qcf_launch_random(prm, l, lg, grid, ctx->stream);
    }
    if (e == QCF_NO_KERNEL) {
        return fail(QCF_ERANGE, "no random kernel for N limbs %d, guess limbs %d", ln, lg);
    }
    QCF_CUDA(e);
    QCF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    QCF_CUDA(cudaMemcpyAsync(ctx->h_hits, ctx->d_hits, QCF_HITS_D2H_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    QCF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    st->launches = 1;
    st->seconds = ms * 1e-3;
    st->counted = ctx->h_hits->tested;
    return QCF_OK;
}

// Chain-draw random mode (QCF_RANDOM_CHAINS): see include/qimcifa_cuda.h for the counter -> guess map.
struct RandomWindow {
    int base_level = 0;
    FilterPlan pl;
    u64 per_group = 0; // guess numbers per row-group sample: 4 rows x n_a chains x K steps
    u32 groups = 0;    // samples per counter block: floor(2^34 / per_group)
};

static int random_window(qcf_ctx* ctx, u128 N, int level, u64 seed, u64 block, RandomWindow* w)
{
    uint32_t n_le[4];
    to_limbs(N, n_le, 4);
    uint64_t node_range = 0, bc = 0;
    uint32_t sqrt_le[4] = { 0, 0, 0, 0 };
    int rc = qcf_node_split(n_le, 4, 1, sqrt_le, nullptr, &node_range, &bc);
    if (rc) {
        return rc;
    }
    {
        // Of every window only the top octave of its value range is drawn (as the sweep cuts its intervals).  The windows of a
        // large number span far less than an octave; for a number so small that the top octave of its one window lies wholly
        // above sqrt(N) (the range is padded to whole batches) chain draws would never test a guess below sqrt(N).
        const u128 top = h_forward(batch_p_hi(bc - 1U));
        if ((top >> 1) + 1U > from_limbs(sqrt_le, 4)) {
            return fail(QCF_ERANGE, "chain-draw random mode needs a number large enough for the filter kernel (sqrt(N) lies below the top octave of the padded range); use flags = 0");
        }
    }
    const u64 n_win = (bc + QCF_RANDOM_WINDOW_BATCHES - 1U) / QCF_RANDOM_WINDOW_BATCHES;
    const u64 wi = qcf_random_window_of(seed, block, n_win);
    const u64 b_hi = bc - wi * QCF_RANDOM_WINDOW_BATCHES, b_lo = (b_hi > QCF_RANDOM_WINDOW_BATCHES) ? (b_hi - QCF_RANDOM_WINDOW_BATCHES) : 0U;
    const u128 v_hi = h_forward(batch_p_hi(b_hi - 1U));
    u128 v_lo = h_forward(batch_p_lo(b_lo));
    if (v_lo < (v_hi >> 1) + 1U) {
        v_lo = (v_hi >> 1) + 1U; // one octave at most, as the sweep cuts its intervals
    }
    w->base_level = best_filter_plan(ctx, N, level, v_lo, v_hi, false, &w->pl);
    if (!w->base_level || (w->pl.tprime < 2U)) {
        return fail(QCF_ERANGE, "chain-draw random mode needs a number large enough for the filter kernel (window %llu of %llu has no plan); use flags = 0",
            (unsigned long long)wi, (unsigned long long)n_win);
    }
    w->per_group = (u64)4U * ctx->tables[w->base_level].n_a * w->pl.steps;
    const u64 per_block = (u64)1 << QCF_RANDOM_BLOCK_LOG2;
    if (w->per_group > per_block) {
        return fail(QCF_ERANGE, "a row group of this window holds more than 2^%d guesses", QCF_RANDOM_BLOCK_LOG2);
    }
    w->groups = (u32)(per_block / w->per_group);
    return QCF_OK;
}

static int run_random_chains(qcf_ctx* ctx, u128 N, int ln, int level, u64 seed, u64 counter_lo, u64 budget, qcf_result* out)
{
    const bool count = level > 5; // chains may run over a lower level than requested (any level above 5): the exact number of level guesses walked is counted on the device
    const u64 per_block = (u64)1 << QCF_RANDOM_BLOCK_LOG2;
    const u128 c_end = (u128)counter_lo + budget;
    if (c_end >> 64) {
        return fail(QCF_EINVAL, "counter_lo + budget exceeds 64 bits");
    }
    QCF_CUDA(cudaMemsetAsync(&ctx->d_hits->count, 0, sizeof(u32), ctx->stream));
    QCF_CUDA(cudaMemsetAsync(&ctx->d_hits->tested, 0, sizeof(u64) + 2 * sizeof(u32) + sizeof(unsigned long long) * QCF_WORK_RING, ctx->stream));
    ctx->ring_next = 0;
    ctx->gcd_mode = false;
    ctx->extra_exact = 0;
    QCF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = lanes_fork(ctx);
    if (rc) {
        return rc;
    }
    u64 launches = 0, tested = 0;
    for (u64 b = counter_lo >> QCF_RANDOM_BLOCK_LOG2; (u128)b * per_block < c_end; ++b) {
        RandomWindow w;
        rc = random_window(ctx, N, level, seed, b, &w);
        if (rc) {
            return rc;
        }
        // the samples of block b whose FIRST guess number lies in [counter_lo, counter_lo + budget)
        const u64 base_ctr = b * per_block;
        const u64 lo_in = (counter_lo > base_ctr) ? (counter_lo - base_ctr) : 0U;
        const u128 hi_in128 = c_end - base_ctr;
        const u64 hi_in = (hi_in128 > per_block) ? per_block : (u64)hi_in128;
        const u64 j_lo = (lo_in + w.per_group - 1U) / w.per_group;
        const u64 j_hi = std::min<u64>(w.groups, (hi_in + w.per_group - 1U) / w.per_group);
        if (j_hi <= j_lo) {
            continue;
        }
        RandomSelection rs;
        rs.seed = seed;
        rs.block = b;
        rs.first = (u32)j_lo;
        rs.count = (u32)(j_hi - j_lo);
        rc = launch_filter_values(ctx, N, ln, ctx->tables[w.base_level], w.pl, count, w.base_level, level, &rs);
        if (rc) {
            return rc;
        }
        ++launches;
        tested += (j_hi - j_lo) * w.per_group;
        ctx->last_plan = w.pl;
        ctx->last_plan_period = ctx->tables[w.base_level].period;
    }
    rc = lanes_join(ctx);
    if (rc) {
        return rc;
    }
    QCF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    QCF_CUDA(cudaMemcpyAsync(ctx->h_hits, ctx->d_hits, QCF_HITS_D2H_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    QCF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    out->kernel_launches += launches;
    out->device_seconds += ms * 1e-3;
    ctx->queue_spills += ctx->h_hits->overflow;
    out->guesses_tested = count ? ctx->h_hits->tested : tested;
    out->guesses_counted = ctx->h_hits->tested;
    return QCF_OK;
}

int qcf_sweep_random(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t budget, uint32_t flags, qcf_result* out)
{
    if (!ctx || !out) {
        return fail(QCF_EINVAL, "ctx/out is null");
    }
    memset(out, 0, sizeof(*out));
    if (!budget) {
        return QCF_OK;
    }
    if (flags & QCF_RANDOM_CHAINS) {
        u128 N;
        int ln;
        int rc = check_n(n_le, n_limbs, &N, &ln);
        if (rc) {
            return rc;
        }
        if ((level < QCF_MIN_LEVEL) || (level > QCF_TABLE_MAX_LEVEL)) {
            return fail(QCF_EINVAL, "random-guess mode draws from the residue tables: levels 5..9");
        }
        QCF_CUDA(cudaSetDevice(ctx->device));
        ctx->opt = qcf_options_snapshot();
        ctx->last_plan_period = 0;
        ctx->queue_spills = 0;
        rc = run_random_chains(ctx, N, ln, level, seed, counter_lo, budget, out);
        if (rc) {
            return rc;
        }
        out->kernel_kind = QCF_KERNEL_FILTER;
        out->aborted = ctx->h_hits->stop ? 1 : 0;
        const u32 nh = std::min<u32>(ctx->h_hits->count, QCF_MAX_HITS);
        out->n_hits = ctx->h_hits->count;
        if (nh) {
            u128 best = from_limbs(ctx->h_hits->g[0], 4);
            for (u32 i = 1; i < nh; ++i) {
                best = std::min(best, from_limbs(ctx->h_hits->g[i], 4));
            }
            out->found = 1;
            to_limbs(best, out->factor_le, 4);
        }
        return QCF_OK;
    }
    RunStats st;
    const int rc = run_random(ctx, n_le, n_limbs, level, seed, counter_lo, budget, nullptr, &st);
    if (rc) {
        return rc;
    }
    out->kernel_launches = st.launches;
    out->device_seconds = st.seconds;
    out->kernel_kind = QCF_KERNEL_EXACT;
    out->guesses_tested = st.counted; // random mode counts on the device (the value 1 is skipped)
    out->guesses_counted = st.counted;
    out->aborted = ctx->h_hits->stop ? 1 : 0;
    const u32 nh = std::min<u32>(ctx->h_hits->count, QCF_MAX_HITS);
    out->n_hits = ctx->h_hits->count;
    if (nh) {
        u128 best = from_limbs(ctx->h_hits->g[0], 4);
        for (u32 i = 1; i < nh; ++i) {
            best = std::min(best, from_limbs(ctx->h_hits->g[i], 4));
        }
        out->found = 1;
        to_limbs(best, out->factor_le, 4);
    }
    return QCF_OK;
}

int qcf_random_guesses(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t count, uint32_t* guesses_le)
{
    if (!ctx || !guesses_le) {
        return fail(QCF_EINVAL, "null argument");
    }
    if (!count) {
        return QCF_OK;
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    u32* d_dump = nullptr;
    QCF_CUDA(cudaMalloc(&d_dump, sizeof(u32) * 4 * count));
    RunStats st;
    int rc = run_random(ctx, n_le, n_limbs, level, seed, counter_lo, count, d_dump, &st);
    if (!rc) {
        cudaError_t e = cudaMemcpy(guesses_le, d_dump, sizeof(u32) * 4 * count, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            rc = fail(QCF_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
        }
    }
    cudaFree(d_dump);
    return rc;
}

int qcf_random_chain_guesses(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, int level, uint64_t seed, uint64_t counter_lo,
    uint64_t count, uint32_t* guesses_le)
{
    if (!ctx || !guesses_le) {
        return fail(QCF_EINVAL, "null argument");
    }
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    if ((level < QCF_MIN_LEVEL) || (level > QCF_TABLE_MAX_LEVEL)) {
        return fail(QCF_EINVAL, "random-guess mode draws from the residue tables: levels 5..9");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    ctx->opt = qcf_options_snapshot();
    const u64 per_block = (u64)1 << QCF_RANDOM_BLOCK_LOG2;
    RandomWindow w;
    u64 have_block = ~0ull;
    std::vector<u32> ta, tb;
    for (u64 i = 0; i < count; ++i) {
        const u64 ctr = counter_lo + i, b = ctr >> QCF_RANDOM_BLOCK_LOG2, c_in = ctr & (per_block - 1U);
        if (b != have_block) {
            rc = random_window(ctx, N, level, seed, b, &w);
            if (rc) {
                return rc;
            }
            have_block = b;
            const QcfTables& t = ctx->tables[w.base_level];
            ta.resize(t.n_a);
            tb.resize(t.n_b);
            QCF_CUDA(cudaMemcpy(ta.data(), t.d_a, sizeof(u32) * t.n_a, cudaMemcpyDeviceToHost));
            QCF_CUDA(cudaMemcpy(tb.data(), t.d_b, sizeof(u32) * t.n_b, cudaMemcpyDeviceToHost));
        }
        const QcfTables& t = ctx->tables[w.base_level];
        u128 g = 0; // 0 = this guess number is skipped
        const u64 j = c_in / w.per_group;
        if (j < w.groups) {
            const u64 n_srows = (((u64)t.n_b << w.pl.tprime) + 3U) / 4U;
            const u64 srow = qcf_random_row_group_of((u32)seed, (u32)(seed >> 32), (u32)b, (u32)(b >> 32), (u32)j, n_srows);
            const u64 within = c_in % w.per_group, e = within / w.pl.steps, k = within % w.pl.steps;
            const u64 row = srow * 4U + e / t.n_a, c = row / t.n_b;
            const u32 r = (u32)(((u64)ta[(size_t)(e % t.n_a)] + tb[(size_t)(row % t.n_b)]) % t.period);
            g = (w.pl.m_lo + c + ((u128)k << w.pl.tprime)) * t.period + r;
            for (int q = w.base_level; q < level; ++q) {
                if (g % PRIMES10[q] == 0U) {
                    g = 0; // not a guess of the requested level: neither tested nor counted
                }
            }
        }
        to_limbs(g, guesses_le + 4 * i, 4);
    }
    return QCF_OK;
}

int qcf_divisible(qcf_ctx* ctx, const uint32_t* n_le, int n_limbs, const uint32_t* g_le, int g_limbs, uint64_t count, uint8_t* out)
{
    if (!ctx || !g_le || !out) {
        return fail(QCF_EINVAL, "null argument");
    }
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    if ((g_limbs < 1) || (g_limbs > 3)) {
        return fail(QCF_EINVAL, "g_limbs must be in [1,3]");
    }
    if (!count) {
        return QCF_OK;
    }
    for (u64 i = 0; i < count; ++i) {
        if (!(g_le[i * g_limbs] & 1U)) {
            return fail(QCF_EINVAL, "guess %llu is even; wheel guesses are odd", (unsigned long long)i);
        }
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    u32 nl[4];
    to_limbs(N, nl, 4);
    u32* d_g = nullptr;
    uint8_t* d_out = nullptr;
    QCF_CUDA(cudaMalloc(&d_g, sizeof(u32) * g_limbs * count));
    QCF_CUDA(cudaMalloc(&d_out, count));
    QCF_CUDA(cudaMemcpyAsync(d_g, g_le, sizeof(u32) * g_limbs * count, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = qcf_launch_divisible(nl, ln, d_g, g_limbs, count, d_out, ctx->stream);
    if (e == cudaSuccess) {
        e = cudaMemcpyAsync(out, d_out, count, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) {
        e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(d_g);
    cudaFree(d_out);
    if (e == QCF_NO_KERNEL) {
        return fail(QCF_ERANGE, "no divisibility kernel for N limbs %d, guess limbs %d", ln, g_limbs);
    }
    QCF_CUDA(e);
    return QCF_OK;
}

int qcf_plan_filter(const uint32_t* n_le, int n_limbs, int level, const uint32_t* v_lo_le, const uint32_t* v_hi_le,
    int forced, qcf_filter_plan* out)
{
    if (!v_lo_le || !v_hi_le || !out) {
        return fail(QCF_EINVAL, "null argument");
    }
    if ((level < QCF_MIN_LEVEL) || (level > QCF_TABLE_MAX_LEVEL)) {
        return fail(QCF_EINVAL, "level %d outside [5,9] (chains run over levels 5..9)", level);
    }
    static_assert(QCF_PLAN_MAX_SEG == QCF_FILTER_MAX_SEG, "plan hook and kernel parameters disagree");
    u128 N;
    int ln;
    int rc = check_n(n_le, n_limbs, &N, &ln);
    if (rc) {
        return rc;
    }
    QcfTables t;
    memset(&t, 0, sizeof(t));
    t.period = 1;
    for (int i = 0; i < level; ++i) {
        t.period *= PRIMES10[i];
    }
    FilterPlan pl;
    memset(out, 0, sizeof(*out));
    out->period = t.period;
    if (!plan_filter(qcf_options_snapshot(), N, t, from_limbs(v_lo_le, 4), from_limbs(v_hi_le, 4), forced != 0, &pl)) {
        return QCF_OK;
    }
    export_plan(pl, t.period, out);
    return QCF_OK;
}

int qcf_last_plan(qcf_ctx* ctx, qcf_filter_plan* out)
{
    if (!ctx || !out) {
        return fail(QCF_EINVAL, "null argument");
    }
    memset(out, 0, sizeof(*out));
    if (ctx->last_plan_period) {
        export_plan(ctx->last_plan, ctx->last_plan_period, out);
    }
    return QCF_OK;
}

int qcf_last_queue_spills(qcf_ctx* ctx, uint64_t* spills)
{
    if (!ctx || !spills) {
        return fail(QCF_EINVAL, "null argument");
    }
    *spills = ctx->queue_spills;
    return QCF_OK;
}

int qcf_io_bytes(uint32_t* h2d_per_chain_launch, uint32_t* h2d_per_row_launch, uint32_t* d2h_per_call)
{
    if (h2d_per_chain_launch) {
        *h2d_per_chain_launch = (uint32_t)sizeof(FilterParams);
    }
    if (h2d_per_row_launch) {
        *h2d_per_row_launch = (uint32_t)sizeof(ExactParams);
    }
    if (d2h_per_call) {
        *d2h_per_call = (uint32_t)QCF_HITS_D2H_BYTES;
    }
    return QCF_OK;
}

int qcf_set_option(const char* name, const char* value)
{
    if (!name) {
        return fail(QCF_EINVAL, "option name is null");
    }
    const QcfOptions dflt;
    const bool reset = !value || !*value;
    std::lock_guard<std::mutex> lock(g_opt_mu);
    int v[4] = { 0, 0, 0, 0 };
    if (!strcmp(name, "xchain_two_stage")) {
        if (!reset && !parse_ints(value, v, 1, 1, 0, 1)) {
            return fail(QCF_EINVAL, "xchain_two_stage takes 0 or 1, got '%s'", value);
        }
        g_opt.xchain_two_stage = reset ? dflt.xchain_two_stage : v[0];
    } else if (!strcmp(name, "chain_min_log2")) {
        if (!reset && !parse_ints(value, v, 1, 1, 16, 40)) {
            return fail(QCF_EINVAL, "chain_min_log2 takes an integer in [16,40], got '%s'", value);
        }
        g_opt.chain_min_log2 = reset ? dflt.chain_min_log2 : v[0];
    } else if (!strcmp(name, "plan_force")) {
        if (!reset && (!parse_ints(value, v, 3, 4, 0, 64) || (v[0] > 3) || (v[1] > 40) || (v[2] > QCF_FILTER_MAX_SEG) || (v[3] > 2))) {
            return fail(QCF_EINVAL, "plan_force takes 'degree(0..3),tprime(0..40),segments(0..%d)[,pad(0..2)]', got '%s'", QCF_FILTER_MAX_SEG, value);
        }
        g_opt.force_d = v[0];
        g_opt.force_tp = v[1];
        g_opt.force_j = v[2];
        g_opt.force_pad = v[3];
    } else if (!strcmp(name, "plan_drift")) {
        if (!reset && !parse_ints(value, v, 1, 1, -1, 2)) {
            return fail(QCF_EINVAL, "plan_drift takes -1 (default), 0, 1 or 2, got '%s'", value);
        }
        g_opt.force_drift = reset ? dflt.force_drift : v[0];
    } else if (!strcmp(name, "plan_packed")) {
        if (!reset && !parse_ints(value, v, 1, 1, 0, 1)) {
            return fail(QCF_EINVAL, "plan_packed takes 0 or 1, got '%s'", value);
        }
        g_opt.force_packed = reset ? dflt.force_packed : v[0];
    } else if (!strcmp(name, "epoch_trips")) {
        if (!reset && !parse_ints(value, v, 1, 1, 0, 1 << 24)) {
            return fail(QCF_EINVAL, "epoch_trips takes an integer in [0,2^24], got '%s'", value);
        }
        g_opt.epoch_trips = reset ? 0u : (unsigned)v[0];
    } else if (!strcmp(name, "debug") || !strcmp(name, "debug_stats")) {
        if (!reset && !parse_ints(value, v, 1, 1, 0, 1)) {
            return fail(QCF_EINVAL, "%s takes 0 or 1, got '%s'", name, value);
        }
        (strcmp(name, "debug") ? g_opt.debug_stats : g_opt.debug) = reset ? 0 : v[0];
    } else if (!strcmp(name, "cost_model")) {
        double c[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        if (!reset) {
            int n = 0;
            const char* p = value;
            while (*p && (n < 8)) {
                char* end = nullptr;
                const double x = strtod(p, &end);
                if ((end == p) || !(x >= 0.0) || (x > 1e9)) {
                    n = -1;
                    break;
                }
                c[n++] = x;
                p = (*end == ',') ? end + 1 : end;
                if ((*end != ',') && *end) {
                    n = -1;
                    break;
                }
            }
            if ((n != 8) || *p) {
                return fail(QCF_EINVAL, "cost_model takes 8 non-negative numbers 'body2,body3,setup2,setup3,segment2,segment3,survivor,uncovered' (0 = fitted value), got '%s'", value);
            }
        }
        for (int i = 0; i < 8; ++i) {
            g_opt.cost[i] = c[i];
        }
    } else {
        return fail(QCF_EINVAL, "unknown option '%s'", name);
    }
    return QCF_OK;
}

int qcf_microbench(qcf_ctx* ctx, int kind, uint64_t iters, double* ops_per_second, double* seconds)
{
    if (!ctx || !ops_per_second) {
        return fail(QCF_EINVAL, "null argument");
    }
    if ((kind < 0) || (kind > 3)) {
        return fail(QCF_EINVAL, "kind must be 0 (IMAD), 1 (IMAD.WIDE), 2 (IADD) or 3 (IMAD.HI)");
    }
    QCF_CUDA(cudaSetDevice(ctx->device));
    const int block = 256, grid = ctx->sm_count * 8;
    QCF_CUDA(qcf_launch_microbench(kind, 16, grid, block, ctx->d_sink, ctx->stream)); // warm-up
    QCF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    QCF_CUDA(qcf_launch_microbench(kind, iters, grid, block, ctx->d_sink, ctx->stream));
    QCF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    QCF_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    QCF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const double ops = (double)iters * 32.0 * (double)block * (double)grid;
    *ops_per_second = ops / (ms * 1e-3);
    if (seconds) {
        *seconds = ms * 1e-3;
    }
    return QCF_OK;
}

} // extern "C"
