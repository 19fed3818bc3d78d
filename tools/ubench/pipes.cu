// Pipe micro-benchmarks for the filter kernel's inner loop (standalone; nvcc -gencode arch=compute_100a,code=sm_100a).
// Each kernel runs a fixed instruction mix per "trip" with independent register operands, 128-thread blocks, 6 blocks per SM
// (the filter kernel's shape); prints SM cycles per trip per warp-scheduler: the pipe's issue interval x instructions.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
typedef uint64_t u64;

template <int KIND> __global__ void __launch_bounds__(128, 6) k(u32 trips, u32* out, const u32* in)
{
    u32 s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = in[i] + threadIdx.x * (2 * i + 1);
    u32 z = in[32] + threadIdx.x, d = in[33] | 1u;
    u32 m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu, acc = 0;
    for (u32 t = 0; t < trips; ++t) {
        if (KIND == 0) { // 32 x VIADDMNMX.U32 (two chains)
#pragma unroll
            for (int i = 0; i < 32; i += 2) { m0 = __viaddmin_u32(z, s[i], m0); m1 = __viaddmin_u32(z, s[i + 1], m1); }
        } else if (KIND == 1) { // 32 x VIADDMNMX.U16x2
#pragma unroll
            for (int i = 0; i < 32; i += 2) { m0 = __viaddmin_u16x2(z, s[i], m0); m1 = __viaddmin_u16x2(z, s[i + 1], m1); }
        } else if (KIND == 2) { // 16 x VIADDMNMX.U16x2 (the packed trip of 32 guesses)
#pragma unroll
            for (int i = 0; i < 16; i += 2) { m0 = __viaddmin_u16x2(z, s[i], m0); m1 = __viaddmin_u16x2(z, s[i + 1], m1); }
        } else if (KIND == 3) { // 16 x VIADDMNMX.U16x2 + 16 x IMAD (independent)
#pragma unroll
            for (int i = 0; i < 16; i += 2) { m0 = __viaddmin_u16x2(z, s[i], m0); m1 = __viaddmin_u16x2(z, s[i + 1], m1); }
#pragma unroll
            for (int i = 16; i < 32; ++i) s[i] = s[i] * d + z;
        } else if (KIND == 4) { // 32 x IMAD
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] = s[i] * d + z;
        } else if (KIND == 5) { // 32 x (IMAD z + u*d) + 16 x VIMNMX3 (what nvcc makes of the degree-1 trip today)
#pragma unroll
            for (int i = 0; i < 32; i += 2) { m0 = min(m0, min(z + (u32)i * d, z + (u32)(i + 1) * d)); }
        } else if (KIND == 6) { // 32 x VIADDMNMX.U32 + 31 IMAD (the degree-2 trip today)
#pragma unroll
            for (int i = 0; i < 32; i += 2) { m0 = __viaddmin_u32(z, s[i], m0); m1 = __viaddmin_u32(z, s[i + 1], m1); }
#pragma unroll
            for (int i = 1; i < 32; ++i) s[i] += (u32)i * d;
        } else if (KIND == 7) { // packed degree 2, G = 2: 16 min-add on base, 16 IMAD, 16 min-add on base + E, 16 packed adds
#pragma unroll
            for (int i = 0; i < 16; i += 2) { m0 = __viaddmin_u16x2(z, s[i], m0); m1 = __viaddmin_u16x2(z, s[i + 1], m1); }
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                u32 x0, x1;
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x0) : "r"(s[16 + i]), "r"(d), "r"(s[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x1) : "r"(s[17 + i]), "r"(d), "r"(s[i + 1]));
                m0 = __viaddmin_u16x2(z, x0, m0);
                m1 = __viaddmin_u16x2(z, x1, m1);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = __vadd2(s[i], s[16 + i]);
        } else if (KIND == 8) { // 32 x PRMT
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] = __byte_perm(s[i], z, 0x7632);
        } else if (KIND == 9) { // 32 x packed add
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] = __vadd2(s[i], z);
        }
        z += d;
        if (KIND != 4 && KIND != 8 && KIND != 9) {
            if (min(m0, m1) == 12345u) { acc += t; m0 = m1 = 0xFFFFFFFFu; } // the rare survivor branch
        }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ m0 ^ m1 ^ z;
}

// semantics of the packed instructions against the portable model the host tests use (qcf_filter_math.cuh)
__global__ void semantics(u32 n, u32* bad)
{
    u32 x = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1u);
    for (u32 i = 0; i < n; ++i) {
        x = x * 1664525u + 1013904223u;
        const u32 a = x;
        x = x * 1664525u + 1013904223u;
        const u32 b = x;
        x = x * 1664525u + 1013904223u;
        const u32 c = (i & 1u) ? x : (x & 0x000F000Fu);
        const u32 lo = (a + b) & 0xFFFFu, hi = ((a >> 16) + (b >> 16)) & 0xFFFFu;
        const u32 want = ((lo < (c & 0xFFFFu)) ? lo : (c & 0xFFFFu)) | (((hi < (c >> 16)) ? hi : (c >> 16)) << 16);
        if (__viaddmin_u16x2(a, b, c) != want) atomicAdd(&bad[0], 1u);
        if (__vadd2(a, b) != (lo | (hi << 16))) atomicAdd(&bad[1], 1u);
        if (__byte_perm(a, b, 0x7632) != ((a >> 16) | (b & 0xFFFF0000u))) atomicAdd(&bad[2], 1u);
    }
}

template <int KIND> void run(const char* name, int sms, int khz, u32* d_out, u32* d_in)
{
    const u32 trips = 200000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<KIND><<<sms * 6, 128>>>(1000, d_out, d_in);
    cudaEventRecord(e0);
    k<KIND><<<sms * 6, 128>>>(trips, d_out, d_in);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM sub-partition: 6 warps walk `trips` trips each
    const double cyc = ms * 1e-3 * khz * 1e3 / (6.0 * trips);
    printf("{\"kind\": %d, \"name\": \"%s\", \"ms\": %.3f, \"cycles_per_trip_per_scheduler\": %.2f}\n", KIND, name, ms, cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    u32 *d_out, *d_in, h_in[34];
    for (int i = 0; i < 34; ++i) h_in[i] = 0x9E3779B9u * (i + 1);
    cudaMalloc(&d_out, sizeof(u32) * p.multiProcessorCount * 6 * 128);
    cudaMalloc(&d_in, sizeof(h_in));
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, khz);
    {
        u32 *d_bad, h_bad[3] = { 0, 0, 0 };
        cudaMalloc(&d_bad, sizeof(h_bad));
        cudaMemset(d_bad, 0, sizeof(h_bad));
        semantics<<<64, 128>>>(4096, d_bad);
        cudaMemcpy(h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost);
        printf("{\"semantics\": \"viaddmin_u16x2 / vadd2 / byte_perm against the portable model\", \"cases\": %d, \"mismatches\": [%u, %u, %u]}\n", 64 * 128 * 4096, h_bad[0], h_bad[1], h_bad[2]);
    }
    run<0>("32 VIADDMNMX.U32", p.multiProcessorCount, khz, d_out, d_in);
    run<1>("32 VIADDMNMX.U16x2", p.multiProcessorCount, khz, d_out, d_in);
    run<2>("16 VIADDMNMX.U16x2", p.multiProcessorCount, khz, d_out, d_in);
    run<3>("16 VIADDMNMX.U16x2 + 16 IMAD", p.multiProcessorCount, khz, d_out, d_in);
    run<4>("32 IMAD", p.multiProcessorCount, khz, d_out, d_in);
    run<5>("32 IMAD + 16 VIMNMX3", p.multiProcessorCount, khz, d_out, d_in);
    run<6>("32 VIADDMNMX.U32 + 31 IMAD", p.multiProcessorCount, khz, d_out, d_in);
    run<7>("packed degree 2, G=2 (two trips)", p.multiProcessorCount, khz, d_out, d_in);
    run<8>("32 PRMT", p.multiProcessorCount, khz, d_out, d_in);
    run<9>("32 packed add", p.multiProcessorCount, khz, d_out, d_in);
    return cudaDeviceSynchronize() != cudaSuccess;
}
