// TEST INFRASTRUCTURE -- NOT a CPU path of the product.  A stand-in for <cuda_runtime.h> that lets g++ compile the
// library's .cu sources (after tests/emu/build_emu.py has rewritten the <<<...>>> launches) into
// tests/_build/libqimcifa_emu.so, where every kernel runs on the host, one block at a time, its threads as fibers
// (ucontext) that meet at __syncthreads() and at warp shuffles.  tests/test_emu_parity.py uses it to run the kernels'
// full control flow (claims, chains, queues, epochs, reports) against the oracle in a container without a GPU, so that
// kernel changes can be checked before GPU time is spent.  The product (qimcifa_b200/abi.py, the host drivers) never
// loads it: libqimcifa_cuda.so has no CPU fallback and qcf_create fails without a device.
#pragma once
#define QCF_EMU 1

#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- host runtime stubs (everything is synchronous) -----------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801 };
typedef struct EmuStream* cudaStream_t;
struct EmuEvent { double t; };
typedef EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDeviceAttr { cudaDevAttrClockRate = 13 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; };
struct cudaIpcMemHandle_t { char reserved[64]; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulator: unsupported"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "host emulator (tests only)");
    p->multiProcessorCount = 2;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 1000000; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{ 0 }; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
    e->t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = (float)std::max(1e-6, (b->t - a->t) * 1e3);
    return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorInvalidValue; }
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// ---- device side: qualifiers, types, intrinsics ---------------------------------------------------------------------
#define __device__
#define __forceinline__ inline
#define __global__
#define __host__
#define __constant__
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static /* one block runs at a time, so block-shared storage is one static object */

struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }

struct EmuThread {
    ucontext_t ctx;
    uint3 tid;
    int state; // 0 runnable, 1 waiting at __syncthreads, 2 waiting at a warp shuffle, 3 done
    unsigned long long shfl_in, shfl_out;
    unsigned shfl_off;
};
extern EmuThread* emu_cur;
extern uint3 emu_blockIdx, emu_blockDim, emu_gridDim;
void* emu_dyn_smem();
void emu_yield(int state);
void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);

#define threadIdx (emu_cur->tid)
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim

static inline void __syncthreads() { emu_yield(1); }
static inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, int off)
{
    emu_cur->shfl_in = v;
    emu_cur->shfl_off = (unsigned)off;
    emu_yield(2);
    return emu_cur->shfl_out;
}
static inline int __any_sync(unsigned, int pred)
{
    emu_cur->shfl_in = pred ? 1ull : 0ull;
    emu_cur->shfl_off = ~0u; // a vote, not a shuffle (emu_runtime.cpp)
    emu_yield(2);
    return emu_cur->shfl_out != 0ull;
}
static inline void __threadfence_system() {}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
// threads of a block are fibers of ONE host thread: plain read-modify-write is atomic with respect to them
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { const unsigned o = *p; *p = o > v ? o : v; return o; }
static inline unsigned atomicExch_system(unsigned* p, unsigned v) { const unsigned o = *p; *p = v; return o; }
static inline unsigned long long atomicAdd_system(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
