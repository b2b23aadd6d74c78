// bench_gather.cu -- diagnostic (not a test): what does one random 4-byte read of a multi-GiB array cost
// on B200, per load flavour?  ncu showed the inverse walk moving 127 B of DRAM traffic per 4-byte step
// (l1tex requests 1 sector, L2 sees 4 from the texture unit): which PTX load avoids the 128-byte fill?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tests/bench_gather tests/bench_gather.cu
//   tests/bench_gather [log2 elements = 30] [gathers per thread = 16]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef uint32_t u32;
typedef uint64_t u64;

enum Flavour { F_PLAIN = 0, F_CG, F_CS, F_CV, F_NC, F_NC_NA, F_NA, F_RELAXED, F_LU, F_EVICT_FIRST, F_NC_L2_64, F_COUNT };
static const char *names[F_COUNT] = {"ld.global (ca)", "ld.global.cg", "ld.global.cs", "ld.global.cv", "ld.global.nc",
                                     "ld.global.nc.L1::no_allocate", "ld.global.L1::no_allocate", "ld.relaxed.gpu.global",
                                     "ld.global.lu", "ld.global.L1::evict_first", "ld.global.nc.L2::64B"};

template <int F>
static __device__ __forceinline__ u32 load(const u32 *p)
{
    u32 v;
    if (F == F_PLAIN) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_CG) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_CS) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_CV) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_NC) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_NA) asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_RELAXED) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_LU) asm volatile("ld.global.lu.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_EVICT_FIRST) asm volatile("ld.global.L1::evict_first.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == F_NC_L2_64) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

static __device__ __forceinline__ u32 mix(u32 x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// independent gathers: G per thread, all issued before the first use
template <int F, int G>
__global__ void __launch_bounds__(256) k_gather(const u32 *__restrict__ a, u32 mask, u32 *__restrict__ out, u32 salt)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 acc = 0;
    u32 v[G];
#pragma unroll
    for (int g = 0; g < G; g++) v[g] = load<F>(a + (mix(t * G + g + salt) & mask));
#pragma unroll
    for (int g = 0; g < G; g++) acc += v[g];
    out[t] = acc;
}

// dependent chase: the loaded value is the next index (array = a random permutation-like map)
template <int F>
__global__ void __launch_bounds__(256) k_chase(const u32 *__restrict__ a, u32 mask, u32 *__restrict__ out, u32 steps)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 i = mix(t) & mask;
    for (u32 s = 0; s < steps; s++) i = load<F>(a + i) & mask;
    out[t] = i;
}

__global__ void k_fill(u32 *a, u64 n)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) a[i] = mix((u32)i * 2654435761u + 12345u);
}

template <int F>
static void run(const u32 *a, u32 mask, u32 *out, u32 threads, int lg)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_g = 0, ms_c = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_gather<F, 16><<<threads / 256, 256>>>(a, mask, out, rep * 7919u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_g, e0, e1);
    }
    const u32 chase_threads = 148 * 2048, steps = 256;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k_chase<F><<<chase_threads / 256, 256>>>(a, mask, out, steps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_c, e0, e1);
    }
    cudaError_t e = cudaGetLastError();
    const double ng = (double)threads * 16, nc = (double)chase_threads * steps;
    printf("%-32s 2^%d u32: gather %7.2f G/s (%6.2f ms)   chase %7.2f G steps/s (%6.2f ms)  %s\n", names[F], lg, ng / ms_g * 1e-6, ms_g,
           nc / ms_c * 1e-6, ms_c, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main(int argc, char **argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 30;
    const u64 n = 1ull << lg;
    const u32 mask = (u32)(n - 1);
    u32 *a, *out;
    const u32 threads = 1u << 24;
    if (cudaMalloc(&a, n * 4) != cudaSuccess || cudaMalloc(&out, (size_t)threads * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    k_fill<<<148 * 8, 256>>>(a, n);
    cudaDeviceSynchronize();
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("array 2^%d u32 = %.1f GiB, cudaLimitMaxL2FetchGranularity = %zu\n", lg, n * 4.0 / (1 << 30), g);
    run<F_PLAIN>(a, mask, out, threads, lg);
    run<F_CG>(a, mask, out, threads, lg);
    run<F_CS>(a, mask, out, threads, lg);
    run<F_CV>(a, mask, out, threads, lg);
    run<F_NC>(a, mask, out, threads, lg);
    run<F_NC_NA>(a, mask, out, threads, lg);
    run<F_NA>(a, mask, out, threads, lg);
    run<F_RELAXED>(a, mask, out, threads, lg);
    run<F_LU>(a, mask, out, threads, lg);
    run<F_EVICT_FIRST>(a, mask, out, threads, lg);
    run<F_NC_L2_64>(a, mask, out, threads, lg);
    if (argc > 3) {
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[3]));
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("-- cudaLimitMaxL2FetchGranularity now %zu\n", g);
        run<F_PLAIN>(a, mask, out, threads, lg);
        run<F_CG>(a, mask, out, threads, lg);
        run<F_NC_NA>(a, mask, out, threads, lg);
    }
    return 0;
}
