// Sweep kernels, table generation and test hooks for sm_100a.
//
// Reference functions replaced here (paths relative to the reference checkout):
//   getSmoothNumbers           include/qimcifa.hpp:384-399   -> qcf_exact_kernel (outer batch loop is on the host)
//   GetWheelIncrement          include/qimcifa.hpp:310-327   -> stateless CRT residue tables (A[i] + B[j]) mod period
//   forward / wheel11          include/qimcifa.hpp:224-259   -> table A of level 5 is wheel11, generated on device
//   getSmoothNumbersIteration  include/qimcifa.hpp:364-372   -> qcf_divides<> (qcf_device.cuh)
//   wheel_inc / wheel_gen      include/qimcifa.hpp:274-308   -> qcf_build_tables_kernel (sieve + scan + scatter)
#include "qcf_chain_count.cuh"
#include "qcf_device.cuh"
#include "qcf_internal.h"

// ---------------------------------------------------------------------------------------------
// Table generation: residues coprime to the wheel primes, on the device.
// ---------------------------------------------------------------------------------------------
__device__ __constant__ u32 c_wheel_primes[9] = { 2, 3, 5, 7, 11, 13, 17, 19, 23 };

// One block.  Marks x in [0, mod) coprime to primes[p_lo..p_hi), block-scans the marks and scatters
// the survivors, ascending, multiplied by `coef` modulo `period`.
__device__ void qcf_sieve_compact(u32 mod, int p_lo, int p_hi, u64 coef, u32 period, u32* out, u32* s_cnt)
{
    const u32 nt = blockDim.x, tid = threadIdx.x;
    const u32 chunk = (mod + nt - 1) / nt;
    const u32 x0 = tid * chunk;
    const u32 x1 = (x0 + chunk < mod) ? (x0 + chunk) : mod;
    u32 cnt = 0;
    for (u32 x = x0; x < x1; ++x) {
        bool ok = true;
        for (int k = p_lo; k < p_hi; ++k) {
            ok = ok && (x % c_wheel_primes[k]);
        }
        cnt += ok ? 1u : 0u;
    }
    s_cnt[tid] = cnt;
    __syncthreads();
    // Hillis-Steele inclusive scan over the per-thread counts
    for (u32 off = 1; off < nt; off <<= 1) {
        const u32 v = (tid >= off) ? s_cnt[tid - off] : 0u;
        __syncthreads();
        s_cnt[tid] += v;
        __syncthreads();
    }
    u32 pos = s_cnt[tid] - cnt;
    for (u32 x = x0; x < x1; ++x) {
        bool ok = true;
        for (int k = p_lo; k < p_hi; ++k) {
            ok = ok && (x % c_wheel_primes[k]);
        }
        if (ok) {
            out[pos++] = (u32)(((u64)x * coef) % period);
        }
    }
    __syncthreads();
}

__device__ u32 qcf_modinv_small(u32 a, u32 m)
{
    // m <= 30030: brute force is fine (runs once per level, one thread)
    a %= m;
    for (u32 x = 1; x < m; ++x) {
        if (((u64)a * x) % m == 1u) {
            return x;
        }
    }
    return 0u; // m == 1
}

__global__ void __launch_bounds__(1024) qcf_build_tables_kernel(int level, u32 period, u32 mod_a, u32 mod_b, u32* out_a, u32* out_b)
{
    __shared__ u32 s_cnt[1024];
    __shared__ u64 s_coef[2];
    const int n_pa = (level < 6) ? level : 6;
    if (threadIdx.x == 0) {
        // CRT coefficients: c1 = 1 (mod mod_a), 0 (mod mod_b); c2 = 0 (mod mod_a), 1 (mod mod_b)
        if (mod_b == 1u) {
            s_coef[0] = 1u;
            s_coef[1] = 0u;
        } else {
            s_coef[0] = (u64)mod_b * qcf_modinv_small(mod_b % mod_a, mod_a);
            s_coef[1] = (u64)mod_a * qcf_modinv_small(mod_a % mod_b, mod_b);
        }
    }
    __syncthreads();
    qcf_sieve_compact(mod_a, 0, n_pa, s_coef[0] % period, period, out_a, s_cnt);
    if (mod_b == 1u) {
        if (threadIdx.x == 0) {
            out_b[0] = 0u;
        }
    } else {
        qcf_sieve_compact(mod_b, 6, level, s_coef[1] % period, period, out_b, s_cnt);
    }
}

cudaError_t qcf_launch_build_tables(int level, QcfTables* t, cudaStream_t stream)
{
    static const u32 primes[9] = { 2, 3, 5, 7, 11, 13, 17, 19, 23 };
    u32 mod_a = 1, mod_b = 1, n_a = 1, n_b = 1;
    for (int k = 0; k < level; ++k) {
        if (k < 6) {
            mod_a *= primes[k];
            n_a *= primes[k] - 1;
        } else {
            mod_b *= primes[k];
            n_b *= primes[k] - 1;
        }
    }
    t->period = mod_a * mod_b;
    t->n_a = n_a;
    t->n_b = n_b;
    cudaError_t e = cudaMalloc(&t->d_a, sizeof(u32) * n_a);
    if (e != cudaSuccess) {
        return e;
    }
    e = cudaMalloc(&t->d_b, sizeof(u32) * n_b);
    if (e != cudaSuccess) {
        return e;
    }
    qcf_build_tables_kernel<<<1, 1024, 0, stream>>>(level, t->period, mod_a, mod_b, t->d_a, t->d_b);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Exact kernel: one guess per thread per step, one Hensel/Montgomery divisibility test per guess.
// Block = 480 threads; a "row" is one (period m, B-residue j) pair = n_a guesses; blocks stride rows.
// The hot loop touches no global memory: N and the bounds sit in registers, the tables in shared memory.
// ---------------------------------------------------------------------------------------------
template <int LG> __device__ __forceinline__ void qcf_load_limbs(u32 (&r)[LG], const u32* src)
{
#pragma unroll
    for (int k = 0; k < LG; ++k) {
        r[k] = src[k];
    }
}

template <int LG> __device__ __forceinline__ void qcf_report_hit(QcfHits* hits, const u32 (&g)[LG])
{
    const u32 idx = qcf_hit_slot(hits);
    if (idx < QCF_MAX_HITS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hits->g[idx][k] = (k < LG) ? g[k] : 0u;
        }
    }
}

template <int LN, int LG, int QN, bool COUNT> __global__ void __launch_bounds__(QCF_BLOCK) qcf_exact_kernel(const ExactParams prm)
{
    extern __shared__ u32 s_tab[];
    u32* s_a = s_tab;
    u32* s_b = s_tab + prm.n_a;
    for (u32 i = threadIdx.x; i < prm.n_a; i += blockDim.x) {
        s_a[i] = prm.d_a[i];
    }
    for (u32 i = threadIdx.x; i < prm.n_b; i += blockDim.x) {
        s_b[i] = prm.d_b[i];
    }
    __syncthreads();

    u32 n[LN];
    qcf_load_limbs<LN>(n, prm.n);
    u32 v_lo[LG], v_hi[LG];
    qcf_load_limbs<LG>(v_lo, prm.v_lo);
    qcf_load_limbs<LG>(v_hi, prm.v_hi);
    const u32 P = prm.period, n_a = prm.n_a, n_b = prm.n_b;

    // row cursor
    u64 row = blockIdx.x;
    u64 moff = row / n_b;
    u32 j = (u32)(row % n_b);
    const u32 dm = gridDim.x / n_b, dj = gridDim.x % n_b;

    // base = prm.base + moff * P ; step = dm * P
    u32 base[LG], step[LG], pl[LG];
    {
        const u64 lo = (u64)(u32)moff * P;
        const u64 hi = (u64)(u32)(moff >> 32) * P + (lo >> 32);
        const u32 prod[3] = { (u32)lo, (u32)hi, (u32)(hi >> 32) };
        const u64 st = (u64)dm * P;
        const u32 stl[3] = { (u32)st, (u32)(st >> 32), 0u };
        u64 c = 0;
#pragma unroll
        for (int k = 0; k < LG; ++k) {
            c += (u64)prm.base[k] + prod[k];
            base[k] = (u32)c;
            c >>= 32;
            step[k] = stl[k];
            pl[k] = (k == 0) ? P : 0u;
        }
    }

    u64 tested = 0;
    u32 iter = 0;
    while (row < prm.n_rows) {
        const u32 bj = s_b[j];
        const bool edge = (moff == 0) || (moff == prm.m_span);
        for (u32 i = threadIdx.x; i < n_a; i += blockDim.x) {
            u32 g0 = s_a[i] + bj;
            g0 = (g0 >= P) ? (g0 - P) : g0;
            u32 g[LG];
            qcf_add32<LG>(g, base, g0);
            bool in = true;
            if (edge) {
                in = !qcf_less<LG>(g, v_lo) && !qcf_less<LG>(v_hi, g);
            }
            if (in) {
                if (COUNT) {
                    if (!(prm.extra_prime && (qcf_mod_small<LG>(g, prm.extra_prime) == 0u))) {
                        ++tested;
                    }
                }
                if (qcf_divides<LN, LG, QN>(n, g)) {
                    if (!(prm.extra_prime && (qcf_mod_small<LG>(g, prm.extra_prime) == 0u))) {
                        qcf_report_hit<LG>(prm.hits, g);
                    }
                }
            }
        }
        row += gridDim.x;
        j += dj;
        moff += dm;
        qcf_addn<LG>(base, step);
        if (j >= n_b) {
            j -= n_b;
            moff += 1;
            qcf_addn<LG>(base, pl);
        }
        if (((++iter) & 63u) == 0u) {
            if (*(volatile u32*)&prm.hits->stop) {
                break;
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            tested += __shfl_down_sync(0xffffffffu, tested, off);
        }
        if ((threadIdx.x & 31u) == 0u) {
            atomicAdd((unsigned long long*)&prm.hits->tested, (unsigned long long)tested);
        }
    }
}

template <int LN, int LG, int QN> static cudaError_t qcf_launch_exact_t(const ExactParams& prm, bool count, int grid, cudaStream_t stream)
{
    const size_t smem = sizeof(u32) * ((size_t)prm.n_a + prm.n_b);
    // attribute once per device (qcf_resident_blocks): a sweep of a small number is a dozen row launches, and the call costs
    // about as much as the launch itself
    static QcfLaunchCache cache[2];
    int per_sm = 0;
    const cudaError_t e = count ? qcf_resident_blocks(qcf_exact_kernel<LN, LG, QN, true>, cache[1], QCF_BLOCK, smem, &per_sm)
                                : qcf_resident_blocks(qcf_exact_kernel<LN, LG, QN, false>, cache[0], QCF_BLOCK, smem, &per_sm);
    if (e != cudaSuccess) {
        return e;
    }
    if (count) {
        qcf_exact_kernel<LN, LG, QN, true><<<grid, QCF_BLOCK, smem, stream>>>(prm);
    } else {
        qcf_exact_kernel<LN, LG, QN, false><<<grid, QCF_BLOCK, smem, stream>>>(prm);
    }
    return cudaGetLastError();
}

cudaError_t qcf_launch_exact(const ExactParams& prm, int ln, int lg, int qn, bool count, int grid, cudaStream_t stream)
{
#define QCF_CASE(LN, LG, QN)                                                                                           \
    if ((ln == LN) && (lg == LG) && (qn == QN)) {                                                                      \
        return qcf_launch_exact_t<LN, LG, QN>(prm, count, grid, stream);                                               \
    }
    // 64-bit N
    QCF_CASE(2, 1, 1) QCF_CASE(2, 1, 2) QCF_CASE(2, 2, 1) QCF_CASE(2, 2, 2)
    // 96-bit N
    QCF_CASE(3, 1, 2) QCF_CASE(3, 1, 3) QCF_CASE(3, 2, 1) QCF_CASE(3, 2, 2) QCF_CASE(3, 2, 3)
    // 128-bit N
    QCF_CASE(4, 1, 3) QCF_CASE(4, 1, 4) QCF_CASE(4, 2, 2) QCF_CASE(4, 2, 3) QCF_CASE(4, 2, 4) QCF_CASE(4, 3, 1)
    QCF_CASE(4, 3, 2) QCF_CASE(4, 3, 3) QCF_CASE(4, 3, 4)
#undef QCF_CASE
    return QCF_NO_KERNEL;
}

// ---------------------------------------------------------------------------------------------
// Test hook: out[i] = g_i | N through the same device routine (full-width quotient, QN = LN).
// ---------------------------------------------------------------------------------------------
template <int LN, int LG> __global__ void qcf_divisible_kernel(const u32 n0, const u32 n1, const u32 n2, const u32 n3, const u32* g_in, u64 count, uint8_t* out)
{
    const u32 nn[4] = { n0, n1, n2, n3 };
    u32 n[LN];
#pragma unroll
    for (int k = 0; k < LN; ++k) {
        n[k] = nn[k];
    }
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
        u32 g[LG];
#pragma unroll
        for (int k = 0; k < LG; ++k) {
            g[k] = g_in[i * LG + k];
        }
        out[i] = qcf_divides<LN, LG, LN>(n, g) ? 1 : 0;
    }
}

cudaError_t qcf_launch_divisible(const u32* n, int ln, const u32* d_g, int lg, u64 count, uint8_t* d_out, cudaStream_t stream)
{
    const int block = 256;
    u64 want = (count + block - 1) / block;
    const int grid = (int)((want < 1) ? 1 : ((want > 148 * 8) ? 148 * 8 : want));
#define QCF_CASE(LN, LG)                                                                                               \
    if ((ln == LN) && (lg == LG)) {                                                                                    \
        qcf_divisible_kernel<LN, LG><<<grid, block, 0, stream>>>(n[0], n[1], n[2], n[3], d_g, count, d_out);           \
        return cudaGetLastError();                                                                                     \
    }
    QCF_CASE(2, 1) QCF_CASE(2, 2) QCF_CASE(3, 1) QCF_CASE(3, 2) QCF_CASE(3, 3) QCF_CASE(4, 1) QCF_CASE(4, 2) QCF_CASE(4, 3)
#undef QCF_CASE
    return QCF_NO_KERNEL;
}

// ---------------------------------------------------------------------------------------------
// Integer-pipe micro-benchmark: 32 independent chains per thread.
// ---------------------------------------------------------------------------------------------
template <int KIND> __global__ void __launch_bounds__(256) qcf_microbench_kernel(u64 iters, u32* sink)
{
    u32 a[32];
    u64 w[16];
    const u32 m = 2654435761u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        a[k] = threadIdx.x * 33u + k;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        w[k] = ((u64)threadIdx.x << 32) | (k + 1u);
    }
    for (u64 it = 0; it < iters; ++it) {
        if (KIND == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(a[(k + 1) & 31]));
            }
        } else if (KIND == 1) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(m), "r"(a[k]));
                }
            }
        } else if (KIND == 2) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(a[(k + 1) & 31]));
            }
        } else { // high half of a 32x32 product plus addend (IMAD.HI.U32): two per guess in the chain kernel's two-stage loop
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(a[(k + 1) & 31]));
            }
        }
    }
    u32 acc = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        acc ^= a[k];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        acc ^= (u32)w[k] ^ (u32)(w[k] >> 32);
    }
    if (acc == 0x12345678u) {
        sink[0] = acc;
    }
}

cudaError_t qcf_launch_microbench(int kind, u64 iters, int grid, int block, u32* d_sink, cudaStream_t stream)
{
    if (kind == 0) {
        qcf_microbench_kernel<0><<<grid, block, 0, stream>>>(iters, d_sink);
    } else if (kind == 1) {
        qcf_microbench_kernel<1><<<grid, block, 0, stream>>>(iters, d_sink);
    } else if (kind == 2) {
        qcf_microbench_kernel<2><<<grid, block, 0, stream>>>(iters, d_sink);
    } else {
        qcf_microbench_kernel<3><<<grid, block, 0, stream>>>(iters, d_sink);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Found flag over peer memory: one thread per peer, a system-scope atomic store into the peer's HBM.
// ---------------------------------------------------------------------------------------------
__global__ void qcf_signal_peers_kernel(u32* const* peer_flags, int n_peers)
{
    // any launch shape reaches every peer (qcf_peer_connect accepts up to 63 of them)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_peers; i += gridDim.x * blockDim.x) {
        atomicExch_system(peer_flags[i], 1u);
    }
    __threadfence_system();
}

// Shared dispenser over peer memory: one system-scope fetch-add on a counter that may live in another GPU's HBM.
__global__ void qcf_dispense_kernel(unsigned long long* word, unsigned long long take, unsigned long long* out)
{
    *out = atomicAdd_system(word, take);
    __threadfence_system();
}

cudaError_t qcf_launch_dispense(unsigned long long* word, unsigned long long take, unsigned long long* out, cudaStream_t stream)
{
    qcf_dispense_kernel<<<1, 1, 0, stream>>>(word, take, out);
    return cudaGetLastError();
}

cudaError_t qcf_launch_signal_peers(u32* const* d_peer_flags, int n_peers, cudaStream_t stream)
{
    if (n_peers <= 0) {
        return cudaSuccess;
    }
    qcf_signal_peers_kernel<<<(n_peers + 31) / 32, 32, 0, stream>>>(d_peer_flags, n_peers);
    return cudaGetLastError();
}
