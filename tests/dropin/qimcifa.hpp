// TEST INFRASTRUCTURE (tests/test_gpu_cli.py::test_reference_program_with_gpu_sweep).  Drop-in demonstration: the UNMODIFIED reference program (reference src/qimcifa.cpp: its main(), dialog, pre-checks,
// wheel_gen call, node split, dispenser globals, worker threads, printSuccess) with its sweep running on the GPU.
//
// This file is found as "qimcifa.hpp" BEFORE the reference's own header (tests/dropin/Makefile puts this directory first
// on the include path).  It includes the real header with its inline sweep function renamed out of the way, then
// defines Qimcifa::getSmoothNumbers() -- the one function INTEGRATION.md replaces (reference include/qimcifa.hpp:384-399)
// -- on top of the C ABI (include/qimcifa_cuda.h).  Nothing of the reference is copied or edited.
//
// Semantics: the intended ones: a worker takes runs of batches from the top of the node's slice
// under the reference's own mutex and globals (batchNumber / batchBound / batchCount, include/qimcifa.hpp:97-100); the
// snapshot's reversed-sentinel loop (D1) lived in the function this file replaces.
#pragma once

#define getSmoothNumbers getSmoothNumbers_reference_cpu
#include_next "qimcifa.hpp"
#undef getSmoothNumbers

#include <atomic>
#include <cstdint>
#include <cstdlib>

extern "C" {
#include "qimcifa_cuda.h"
}

namespace Qimcifa {

inline void dropinToLimbs(const BigInteger& v, uint32_t* le, int n)
{
    BigInteger t = v;
    for (int i = 0; i < n; ++i) {
        le[i] = (uint32_t)(size_t)(t & 0xFFFFFFFFU);
        t >>= 32U;
    }
}

inline BigInteger dropinFromLimbs(const uint32_t* le, int n)
{
    BigInteger v = 0U;
    for (int i = n - 1; i >= 0; --i) {
        v = v * 4294967296ULL + le[i];
    }
    return v;
}

// The reference starts hardware_concurrency() workers (src/qimcifa.cpp:66,172-178); here a worker is a GPU, so the
// first <number of GPUs> workers to arrive take one GPU each and the rest have nothing to do.
inline bool getSmoothNumbers(const BigInteger& toFactor, std::vector<boost::dynamic_bitset<uint64_t>>& inc_seqs,
    const BigInteger& /* offset: the constant 1 of the default build, src/qimcifa.cpp:148 */,
    const std::chrono::time_point<std::chrono::high_resolution_clock>& iterClock)
{
    static std::atomic<int> nextGpu(0);
    const int gpu = nextGpu.fetch_add(1);
    qcf_ctx* ctx = nullptr;
    if (qcf_create(gpu, &ctx) != QCF_OK) {
        if (gpu == 0) { // no GPU at all: the library has no CPU fallback, so say so instead of sweeping nothing quietly
            std::cerr << "qcf_create failed: " << qcf_last_error() << std::endl;
        }
        return false; // otherwise a surplus worker: every GPU already has one
    }
    uint32_t n_le[QCF_MAX_N_LIMBS];
    dropinToLimbs(toFactor, n_le, QCF_MAX_N_LIMBS);
    const int level = (int)inc_seqs.size() + MIN_RTD_LEVEL; // src/qimcifa.cpp:152-153
    const char* e = getenv("QCF_DROPIN_BATCHES");
    const uint64_t want = e ? strtoull(e, nullptr, 10) : 1024U;
    bool found = false;
    for (;;) {
        uint64_t lo, hi;
        {
            std::lock_guard<std::mutex> lock(batchMutex);
            if (batchNumber >= batchBound) {
                break; // exhausted, or finish() was called (include/qimcifa.hpp:102-105)
            }
            const BigInteger left = batchBound - batchNumber;
            const uint64_t take = (left < BigInteger(want)) ? (uint64_t)(size_t)left : want;
            hi = (uint64_t)(size_t)(batchCount - batchNumber); // node k owns the k-th slice from the top, swept descending
            lo = hi - take;
            batchNumber += take;
        }
        qcf_result res;
        if (qcf_sweep(ctx, n_le, QCF_MAX_N_LIMBS, level, lo, hi, want, QCF_KERNEL_AUTO, &res) != QCF_OK) {
            std::cerr << "qcf_sweep failed: " << qcf_last_error() << std::endl;
            finish();
            break;
        }
        if (res.found) {
            const BigInteger base = dropinFromLimbs(res.factor_le, QCF_MAX_N_LIMBS);
            printSuccess(base, toFactor / base, toFactor, "Exact factor: Found ", iterClock); // include/qimcifa.hpp:370
            found = true;
            break;
        }
    }
    qcf_destroy(ctx);
    return found;
}

} // namespace Qimcifa
