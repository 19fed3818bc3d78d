// Host side of the qimcifa drivers on top of the C ABI (include/qimcifa_cuda.h).
//
// Mirrors the reference's algorithm header (reference include/qimcifa.hpp, namespace Qimcifa) for the
// sweep path: same names, same argument meaning, same output lines -- but getSmoothNumbers() hands batch
// ranges to the GPU instead of looping over guesses, and the dispenser implements the INTENDED split
// (node k owns the k-th slice from the top; the reversed-sentinel defect of the snapshot, SURVEY.md D1,
// is not reproduced).  The prompts and result lines are the reference's CLI contract and are reproduced byte for byte
// (each cites the line it must match; the reference is LGPL-3.0); the code around them is this repository's own.
#pragma once

#include <algorithm>
#include <chrono>
#include <iostream>
#include <mutex>
#include <string>
#include <vector>

#include "big_integer.hpp"
#include "qimcifa_cuda.h"

namespace Qimcifa {

constexpr int BIGGEST_WHEEL = 223092870; // reference include/qimcifa.hpp:86
constexpr int SMALLEST_WHEEL = 2310;     // :88
constexpr int MIN_RTD_LEVEL = 5;         // :89

// ---- dispenser (reference include/qimcifa.hpp:97-129) ------------------------------------------------
// batchNumber counts batches handed out by this node; the batch returned is batchCount-1-batchNumber,
// i.e. batches descend from the top of the node's slice.  finish() exhausts the dispenser.
inline BigInteger batchNumber;
inline BigInteger batchBound;
inline BigInteger batchCount;
inline std::mutex batchMutex;
inline std::vector<qcf_ctx*> activeContexts; // GPUs to stop when a factor is found

inline void finish()
{
    std::lock_guard<std::mutex> lock(batchMutex);
    batchNumber = batchBound;
    for (qcf_ctx* c : activeContexts) {
        qcf_set_found(c, 1); // seen by kernels in flight, like peers observing the exhausted dispenser
    }
}

// Hands out up to `want` consecutive batches [lo, hi) from the top of what is left; false when exhausted.
inline bool getNextBatches(uint64_t want, uint64_t& lo, uint64_t& hi)
{
    std::lock_guard<std::mutex> lock(batchMutex);
    if (batchNumber >= batchBound) {
        return false;
    }
    BigInteger left = batchBound - batchNumber;
    const uint64_t take = (left < BigInteger(want)) ? (uint64_t)left : want;
    hi = (uint64_t)(batchCount - batchNumber);
    lo = hi - take;
    batchNumber += take;
    return true;
}

// Single-batch form with the reference's signature (include/qimcifa.hpp:107-129): returns batchBound as the
// "exhausted" sentinel.  Callers must test exhaustion with isExhausted(), not by comparing.
inline BigInteger getNextBatch()
{
    uint64_t lo, hi;
    if (!getNextBatches(1, lo, hi)) {
        return batchBound;
    }
    return BigInteger(lo);
}

inline bool isPowerOfTwo(const BigInteger& x) { return (bool)x && !(x & (x - 1ULL)); } // :188-199

// floor(sqrt(n)) -- what the reference's bisection (include/qimcifa.hpp:162-186) returns -- by Newton's iteration from above:
// x0 = 2^ceil(bits/2) >= sqrt(n), x <- (x + n/x)/2 decreases strictly until it reaches the floor.
inline BigInteger sqrt(const BigInteger& toTest)
{
    if (toTest < 2U) {
        return toTest;
    }
    BigInteger x = BigInteger(1U) << (unsigned)((toTest.bitLength() + 1) / 2);
    for (;;) {
        const BigInteger next = (x + toTest / x) >> 1U;
        if (next >= x) {
            return x;
        }
        x = next;
    }
}

struct Wheel11Table {
    unsigned short v[480];
    Wheel11Table()
    {
        int k = 0;
        for (int i = 1; i < SMALLEST_WHEEL; ++i) {
            if ((i % 2) && (i % 3) && (i % 5) && (i % 7) && (i % 11)) {
                v[k++] = (unsigned short)i;
            }
        }
    }
};
inline const Wheel11Table wheel11;

inline BigInteger forward(const BigInteger& p) { return wheel11.v[(size_t)(p % 480U)] + (p / 480U) * 2310U; } // :256-259
inline BigInteger backward(const BigInteger& n)                                                                  // :261-263
{
    const unsigned short r = (unsigned short)(size_t)(n % 2310U);
    return (size_t)(std::lower_bound(wheel11.v, wheel11.v + 480, r) - wheel11.v) + 480U * (n / 2310U) + 1U;
}

// The three result lines of the reference (include/qimcifa.hpp:212-222), byte for byte: the factorisation, the time since
// the end of the dialog in seconds at microsecond resolution, and the join notice.  Stops every worker first.
using Clock = std::chrono::high_resolution_clock;
inline void printSuccess(const BigInteger& f1, const BigInteger& f2, const BigInteger& toFactor, const std::string& message,
    const Clock::time_point& iterClock)
{
    finish();
    const long long micros = std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - iterClock).count();
    std::cout << message << f1 << " * " << f2 << " = " << toFactor << std::endl
              << "(Time elapsed: " << (micros / 1000000.0) << " seconds)" << std::endl
              << "(Waiting to join other threads...)" << std::endl;
}

struct SweepTotals {
    uint64_t guesses = 0, batches = 0, launches = 0;
    double deviceSeconds = 0;
};

// The sweep (reference include/qimcifa.hpp:384-399).  One caller per GPU context, the way the reference runs
// one caller per CPU thread (src/qimcifa.cpp:160-178): pull batch ranges from the shared dispenser, sweep
// them on the GPU, report the factor.  `level` replaces inc_seqs (its size + 5 was the level).
inline bool getSmoothNumbers(qcf_ctx* ctx, const BigInteger& toFactor, int level, uint64_t batchesPerLaunch, uint32_t flags,
    const std::chrono::time_point<std::chrono::high_resolution_clock>& iterClock, SweepTotals* totals = nullptr)
{
    uint32_t n_le[QCF_MAX_N_LIMBS];
    toFactor.toLimbs(n_le, QCF_MAX_N_LIMBS);
    uint64_t lo, hi;
    while (getNextBatches(batchesPerLaunch, lo, hi)) {
        qcf_result res;
        const int rc = qcf_sweep(ctx, n_le, QCF_MAX_N_LIMBS, level, lo, hi, batchesPerLaunch, flags, &res);
        if (rc) {
            std::cerr << "qcf_sweep failed: " << qcf_last_error() << std::endl;
            finish();
            return false;
        }
        if (totals) {
            totals->guesses += res.guesses_tested;
            totals->batches += res.batches_done;
            totals->launches += res.kernel_launches;
            totals->deviceSeconds += res.device_seconds;
        }
        if (res.found) {
            const BigInteger base = BigInteger::fromLimbs(res.factor_le, QCF_MAX_N_LIMBS);
            // :369-371, or the common-factor build's line :374-376 (factor_le is gcd(guess, N) in that mode)
            printSuccess(base, toFactor / base, toFactor, (flags & QCF_MODE_GCD) ? "Has common factor: Found " : "Exact factor: Found ", iterClock);
            return true;
        }
        if (res.aborted) {
            return false;
        }
    }
    return false;
}

} // namespace Qimcifa
