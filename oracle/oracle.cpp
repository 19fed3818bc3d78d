// CPU oracle (TEST INFRASTRUCTURE ONLY) -- a plain C++17 restatement of the
// reverse-wheel trial-division sweep of vm6502q/qimcifa for N < 2^128.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product path
// (qimcifa_b200/, libqimcifa_cuda.so) never links, loads or calls it.
//
// Reference lines restated (paths relative to the reference checkout):
//   constants                    include/qimcifa.hpp:86-89
//   dispenser                    include/qimcifa.hpp:97-129
//   floor sqrt                   include/qimcifa.hpp:162-186
//   wheel11 / forward / backward include/qimcifa.hpp:224-263
//   wheel bitsets                include/qimcifa.hpp:265-327
//   divisibility test + sweep    include/qimcifa.hpp:364-399
//   gcd / common-factor mode     include/qimcifa.hpp:201-210,373-379 (IS_RSA_SEMIPRIME=OFF)
//   range / node split           src/qimcifa.cpp:128-158
//   tuner cardinality            src/qimcifa_tuner.cpp:132-151
//
// Arithmetic: the reference's default BigInteger is boost::multiprecision::cpp_int
// (include/qimcifa.hpp:77-78,94; third party, NOT vendored, no version pinned).
// Every operation on the path is exact integer arithmetic, so `unsigned __int128`
// is an exact stand-in for N < 2^128.
//
// Parity status: the reference has no tests or golden vectors.  This oracle is
// pinned against (a) the reference's embedded wheel11 table
// (tests/golden/wheel11.json), (b) the Python-int twin oracle/oracle_py.py, and
// (c) traces of the UNMODIFIED reference sources compiled with compat shims
// (oracle/_ref/qimcifa_ref, tests/golden/ref_traces.json).
//
// All 128-bit values cross the C interface as uint64_t[2] = {lo, hi}.

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

namespace {

constexpr u64 BW_DEFAULT = 223092870ULL; // include/qimcifa.hpp:86
constexpr unsigned PRIMES[10] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29 }; // src/qimcifa.cpp:64

struct Wheel11 {
    uint16_t w[480];
    Wheel11()
    {
        // include/qimcifa.hpp:224-254 (same set src/wheel_gen.py:41-43 prints)
        int k = 0;
        for (int i = 1; i < 2310; ++i) {
            if ((i % 2) && (i % 3) && (i % 5) && (i % 7) && (i % 11)) {
                w[k++] = (uint16_t)i;
            }
        }
    }
};
const Wheel11 W11;

inline u128 ld(const u64* p) { return ((u128)p[1] << 64) | p[0]; }
inline void st(u64* p, u128 v)
{
    p[0] = (u64)v;
    p[1] = (u64)(v >> 64);
}

// include/qimcifa.hpp:256-259
inline u128 forward(u128 p) { return (u128)W11.w[(size_t)(p % 480U)] + (p / 480U) * 2310U; }

// include/qimcifa.hpp:261-263
inline u128 backward(u128 n)
{
    const unsigned r = (unsigned)(n % 2310U);
    unsigned lb = 0;
    while ((lb < 480U) && (W11.w[lb] < r)) {
        ++lb;
    }
    return (u128)lb + (u128)480U * (n / 2310U) + 1U;
}

// include/qimcifa.hpp:162-186.  The reference bisects on [1, n>>1] with an
// arbitrary-precision mid*mid; floor(sqrt(n)) < 2^64 for n < 2^128, so clamping
// the upper end keeps mid*mid inside 128 bits and returns the same value.
inline u128 isqrt(u128 n)
{
    u128 start = 1U, end = n >> 1U, ans = 0U;
    const u128 cap = ((u128)1 << 64) - 1U;
    if (end > cap) {
        end = cap;
    }
    if (start > end) {
        return ans; // n in {0,1,2,3}: do { } while runs once in the reference
    }
    do {
        const u128 mid = (start + end) >> 1U;
        const u128 sqr = mid * mid;
        if (sqr == n) {
            return mid;
        }
        if (sqr < n) {
            start = mid + 1U;
            ans = mid;
        } else {
            end = mid - 1U;
        }
    } while (start <= end);
    return ans;
}

inline bool coprime_level(u128 g, int level)
{
    for (int i = 5; i < level; ++i) {
        if ((g % PRIMES[i]) == 0) {
            return false;
        }
    }
    return true;
}

u128 count_coprime_upto(u128 x, int nprimes)
{
    // #{1<=v<=x coprime to the first nprimes primes}, inclusion-exclusion
    if (x == 0) {
        return 0;
    }
    __int128 total = 0;
    for (unsigned mask = 0; mask < (1U << nprimes); ++mask) {
        u128 d = 1;
        int bits = 0;
        for (int i = 0; i < nprimes; ++i) {
            if ((mask >> i) & 1U) {
                d *= PRIMES[i];
                ++bits;
            }
        }
        if (bits & 1) {
            total -= (__int128)(x / d);
        } else {
            total += (__int128)(x / d);
        }
    }
    return (u128)total;
}

// ---- literal wheel state (include/qimcifa.hpp:274-327) -------------------
// A rotate-right-by-one of a bitset is a head-pointer increment on a circular
// buffer; when bit0 is clear the reference shifts a zero in at the top, which
// is the same rotation.
struct LiteralWheel {
    std::vector<uint8_t> bits; // already shifted by one (:295)
    size_t head = 0;
};

LiteralWheel wheel_inc(int nprimes, u128 limit)
{
    u128 radius = 1U;
    for (int i = 0; i < nprimes; ++i) {
        radius *= PRIMES[i];
    }
    if (limit < radius) {
        radius = limit;
    }
    const unsigned prime = PRIMES[nprimes - 1];
    std::vector<uint8_t> o;
    for (u64 i = 1U; i <= (u64)radius; ++i) {
        bool mult = false;
        for (int k = 0; k < nprimes - 1; ++k) {
            if ((i % PRIMES[k]) == 0) {
                mult = true;
                break;
            }
        }
        if (!mult) {
            o.push_back((i % prime) == 0);
        }
    }
    LiteralWheel w;
    w.bits.assign(o.size(), 0);
    // output >>= 1U: bit j <- bit j+1, top bit cleared
    for (size_t j = 0; j + 1 < o.size(); ++j) {
        w.bits[j] = o[j + 1];
    }
    return w;
}

struct LiteralWheels {
    std::vector<LiteralWheel> seqs;
    LiteralWheels(int level, u128 n)
    {
        // wheel_gen builds one per prefix and the caller erases the first 5
        // (include/qimcifa.hpp:300-308, src/qimcifa.cpp:152-153)
        for (int np = 6; np <= level; ++np) {
            seqs.push_back(wheel_inc(np, n));
        }
    }
    size_t increment()
    {
        size_t inc = 0U;
        bool hit = false;
        do {
            hit = false;
            for (auto& w : seqs) {
                hit = w.bits[w.head] != 0;
                // rotate right by one
                w.bits[w.head] = hit ? 1 : 0; // the bit that left re-enters at the top iff it was set;
                                              // a clear bit re-entering as clear is the plain shift
                w.head = (w.head + 1U == w.bits.size()) ? 0U : (w.head + 1U);
                if (hit) {
                    break;
                }
            }
            ++inc;
        } while (hit);
        return inc;
    }
};

struct LiteralDispenser { // include/qimcifa.hpp:97-129, src/qimcifa.cpp:156-158
    u128 batchNumber, batchBound, batchCount;
    LiteralDispenser(u128 nodeRange, u64 nodeCount, u64 nodeId)
        : batchNumber((u128)nodeId * nodeRange)
        , batchBound((u128)(nodeId + 1U) * nodeRange)
        , batchCount((u128)nodeCount * nodeRange)
    {
    }
    u128 next()
    {
        const u128 result = batchCount - (batchNumber + 1U);
        if (batchNumber >= batchBound) {
            return batchBound;
        }
        ++batchNumber;
        return result;
    }
};

} // namespace

extern "C" {

void orc_wheel11(uint16_t* out) { memcpy(out, W11.w, sizeof(W11.w)); }

void orc_forward(const u64* p, u64* out) { st(out, forward(ld(p))); }
void orc_backward(const u64* n, u64* out) { st(out, backward(ld(n))); }
void orc_isqrt(const u64* n, u64* out) { st(out, isqrt(ld(n))); }

// src/qimcifa.cpp:128,149,155-158
void orc_node_split(const u64* n, u64 node_count, u64 bw, u64* s, u64* full_range, u64* node_range, u64* batch_count)
{
    const u128 N = ld(n);
    const u128 fullMaxBase = isqrt(N);
    const u128 fullRange = backward(fullMaxBase);
    const u128 nodeRange = (((fullRange + node_count - 1U) / node_count) + bw - 1U) / bw;
    st(s, fullMaxBase);
    st(full_range, fullRange);
    st(node_range, nodeRange);
    st(batch_count, (u128)node_count * nodeRange);
}

// include/qimcifa.hpp:201-210
inline u128 ref_gcd(u128 n1, u128 n2)
{
    while (n2 != 0U) {
        const u128 t = n1;
        n1 = n2;
        n2 = t % n2;
    }
    return n1;
}

// What getSmoothNumbersIteration reports for one guess (include/qimcifa.hpp:364-381): mode 0 = IS_RSA_SEMIPRIME
// (exact divisor), mode 1 = the common-factor build (gcd(base, toFactor) != 1).
inline bool guess_hits(u128 N, u128 g, int mode)
{
    return mode ? (ref_gcd(g, N) != 1U) : ((N % g) == 0U);
}

int orc_intent_sweep_indices_mode(const u64* n, int level, const u64* p_lo, const u64* p_hi, int mode, u64* tested, u64* hits, int max_hits);

// S-intent over an inclusive index interval; hits ascending (up to max_hits kept).
int orc_intent_sweep_indices(const u64* n, int level, const u64* p_lo, const u64* p_hi, u64* tested, u64* hits, int max_hits)
{
    return orc_intent_sweep_indices_mode(n, level, p_lo, p_hi, 0, tested, hits, max_hits);
}

int orc_intent_sweep_indices_mode(const u64* n, int level, const u64* p_lo, const u64* p_hi, int mode, u64* tested, u64* hits, int max_hits)
{
    const u128 N = ld(n);
    const u128 lo = ld(p_lo), hi = ld(p_hi);
    u64 cnt = 0;
    int nh = 0;
    if (hi >= lo) {
        u128 k = lo / 480U;
        unsigned r = (unsigned)(lo % 480U);
        u128 base = k * 2310U;
        for (u128 p = lo;; ++p) {
            const u128 g = base + W11.w[r];
            if (coprime_level(g, level)) {
                ++cnt;
                if (guess_hits(N, g, mode)) { // include/qimcifa.hpp:369 / :374-375
                    if (nh < max_hits) {
                        st(hits + 2 * nh, g);
                    }
                    ++nh;
                }
            }
            if (p == hi) {
                break;
            }
            if (++r == 480U) {
                r = 0;
                base += 2310U;
            }
        }
    }
    *tested = cnt;
    return nh;
}

void orc_intent_count(int level, const u64* p_lo, const u64* p_hi, u64* out)
{
    const u128 lo = ld(p_lo), hi = ld(p_hi);
    if (hi < lo) {
        st(out, 0);
        return;
    }
    st(out, count_coprime_upto(forward(hi), level) - count_coprime_upto(forward(lo) - 1U, level));
}

// src/qimcifa_tuner.cpp:133,137-151: the `cardinality` column for levels 5..9
void orc_tuner_cardinalities(const u64* n, u64* out /* 5 x {lo,hi} */)
{
    u128 range = backward(isqrt(ld(n)));
    for (int i = 5; i < 10; ++i) {
        st(out + 2 * (i - 5), range);
        if (i < 9) {
            const unsigned q = PRIMES[i];
            range *= q - 1U;
            range = (range + q - 1U) / q;
        }
    }
}

// src/qimcifa_tuner.cpp:132
void orc_tuner_batch_count(const u64* n, u64 bw, u64* out) { st(out, (isqrt(ld(n)) + bw - 1U) / bw); }

// Batches one thread visits under the literal control flow (D1).
int orc_literal_batches(u64 node_range, u64 node_count, u64 node_id, u64* out, int cap)
{
    LiteralDispenser d(node_range, node_count, node_id);
    int k = 0;
    for (u128 b = d.next(); b < d.batchBound; b = d.next()) {
        if (k < cap) {
            out[k] = (u64)b;
        }
        ++k;
    }
    return k;
}

// Single-thread literal sweep (include/qimcifa.hpp:384-399).  Returns 1 when a factor
// was found.  trace (optional) receives the first trace_cap tested guesses as {lo,hi}.
// A 64-bit FNV-1a hash over every tested guess (16 bytes LE each) is written to *hash.
int orc_literal_sweep_mode(const u64* n, int level, u64 node_count, u64 node_id, u64 bw, u64 max_guesses, int mode, u64* trace,
    u64 trace_cap, u64* tested_out, u64* factor, u64* hash);

int orc_literal_sweep(const u64* n, int level, u64 node_count, u64 node_id, u64 bw, u64 max_guesses, u64* trace,
    u64 trace_cap, u64* tested_out, u64* factor, u64* hash)
{
    return orc_literal_sweep_mode(n, level, node_count, node_id, bw, max_guesses, 0, trace, trace_cap, tested_out, factor, hash);
}

// mode 1: the common-factor build; *factor receives gcd(guess, N), what printSuccess prints first (include/qimcifa.hpp:374-376)
int orc_literal_sweep_mode(const u64* n, int level, u64 node_count, u64 node_id, u64 bw, u64 max_guesses, int mode, u64* trace,
    u64 trace_cap, u64* tested_out, u64* factor, u64* hash)
{
    const u128 N = ld(n);
    const u128 fullRange = backward(isqrt(N));
    const u128 nodeRange = (((fullRange + node_count - 1U) / node_count) + bw - 1U) / bw;
    LiteralWheels wheels(level, N);
    LiteralDispenser disp(nodeRange, node_count, node_id);
    u64 tested = 0;
    u64 h = 1469598103934665603ULL;
    int found = 0;
    for (u128 b = disp.next(); (b < disp.batchBound) && !found; b = disp.next()) {
        const u128 batchStart = b * bw + 1U;
        const u128 batchEnd = (b + 1U) * bw + 1U;
        for (u128 p = batchStart; p < batchEnd;) {
            p += wheels.seqs.empty() ? 1U : wheels.increment();
            const u128 g = forward(p);
            if (tested < trace_cap) {
                st(trace + 2 * tested, g);
            }
            ++tested;
            for (int i = 0; i < 16; ++i) {
                h ^= (u64)((g >> (8 * i)) & 0xFF);
                h *= 1099511628211ULL;
            }
            if (guess_hits(N, g, mode)) {
                st(factor, mode ? ref_gcd(g, N) : g);
                found = 1;
                break;
            }
            if (max_guesses && (tested >= max_guesses)) {
                *tested_out = tested;
                *hash = h;
                return 0;
            }
        }
    }
    *tested_out = tested;
    *hash = h;
    return found;
}

// Multithreaded S-intent sweep of batches [batch_lo, batch_hi) handed out DESCENDING from a
// mutex-guarded counter, as the reference's workers do (src/qimcifa.cpp:160-178,
// include/qimcifa.hpp:107-129).  Per guess it does what the reference does per guess:
// forward(p) with a division by 480 and N % g (include/qimcifa.hpp:258,369).  Workers
// observe the found flag only between batches, like the reference.  This is the "port"
// CPU baseline.  Returns 1 when a factor was found.
int orc_mt_sweep(const u64* n, int level, u64 batch_lo, u64 batch_hi, u64 bw, unsigned threads, u64* tested_out,
    double* seconds, u64* factor)
{
    const u128 N = ld(n);
    std::mutex m;
    u64 next = batch_hi;
    std::atomic<int> found(0);
    std::atomic<u64> tested(0);
    u128 fac = 0;
    auto worker = [&]() {
        for (;;) {
            u64 b;
            {
                std::lock_guard<std::mutex> lock(m);
                if (found.load() || (next <= batch_lo)) {
                    return;
                }
                b = --next;
            }
            u64 cnt = 0;
            const u128 end = ((u128)b + 1U) * bw + 1U;
            for (u128 p = (u128)b * bw + 2U; p <= end; ++p) {
                const u128 g = forward(p);
                if (!coprime_level(g, level)) {
                    continue;
                }
                ++cnt;
                if ((N % g) == 0U) {
                    std::lock_guard<std::mutex> lock(m);
                    if (!found.load() || (g < fac)) {
                        fac = g;
                    }
                    found.store(1);
                }
            }
            tested += cnt;
        }
    };
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) {
        pool.emplace_back(worker);
    }
    for (auto& t : pool) {
        t.join();
    }
    *seconds = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count() * 1e-9;
    *tested_out = tested.load();
    st(factor, fac);
    return found.load();
}

} // extern "C"
