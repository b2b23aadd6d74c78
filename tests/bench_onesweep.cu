// bench_onesweep.cu -- micro-benchmark of the onesweep radix pass (experiments, not a test).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DOS_PROFILE_PHASES] \
//        -o tests/bench_onesweep tests/bench_onesweep.cu
//   tests/bench_onesweep [log2_m=26] [reps=5]
//
// For every kernel configuration: P passes over m (u64 key, u32 value) pairs, ping-pong, each pass
// timed with CUDA events; after the first pass a checker kernel verifies that the output is the
// stable partition by digit of the input (digit order, value order inside a digit, key/value
// pairing, checksum).  Key distributions: uniform random digits and "text-like" skewed digits.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#include "../bijective-bwt_b200/csrc/radix.cuh"

#define CHECK(x)                                                                                  \
    do {                                                                                          \
        cudaError_t e = (x);                                                                      \
        if (e != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e));   \
            exit(2);                                                                              \
        }                                                                                         \
    } while (0)

static __device__ __forceinline__ u64 mix64(u64 x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// dist 0: uniform; dist 1: every byte is the AND of two uniform bytes OR-ed with a third AND-ed pair
// (skewed digits: a few very common values, a long tail)
__global__ void k_gen(u64 *keys, u32 m, int dist, u64 seed)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u64 a = mix64(seed + i);
    if (dist == 1) {
        const u64 b = mix64(a), c = mix64(b), d = mix64(c);
        a = (a & b & c) | (b & c & d & mix64(d));
    }
    keys[i] = a;
}

// pairing: the value carried with a key is the index the key had in the input of this pass chain
__global__ void k_check(const u64 *kin, const u64 *kout, const u32 *vout, u32 m, u32 shift, u32 *errors)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u64 k = kout[i];
    const u32 v = vout[i];
    bool bad = v >= m || kin[v] != k;
    if (i + 1 < m) {
        const u32 d0 = (u32)(k >> shift) & 255, d1 = (u32)(kout[i + 1] >> shift) & 255;
        if (d0 > d1) bad = true;
        if (d0 == d1 && v >= vout[i + 1]) bad = true;
    }
    if (bad) atomicAdd(errors, 1u);
}

struct Bufs {
    u64 *k[2];
    u32 *v[2];
    u32 *hist;
    u64 *status;
    u32 *errors;
    u64 *orig;
};

static u32 g_epoch = 0;

typedef void (*os_kernel_t)(const u64 *, const u32 *, u64 *, u32 *, u32, u32, const u32 *, u64 *, u32);
// launch(kin, vin, kout, vout, pass)
typedef std::function<void(const u64 *, const u32 *, u64 *, u32 *, int)> launcher_t;

static void run_kernel(const char *name, const void *kern, launcher_t launch, int NT, int IPT, size_t smem_bytes, Bufs &b,
                       u32 m, int reps, int dist, launcher_t pre = nullptr);

template <int NT, int IPT, int MINB, int LB>
static void run_config(const char *name, Bufs &b, u32 m, int reps, int dist)
{
    os_kernel_t kern = k_onesweep_pass<u64, NT, IPT, MINB, LB>;
    const size_t sm = OsSmem<u64, NT, IPT>::bytes;
    run_kernel(name, (const void *)kern, [=](const u64 *ki, const u32 *vi, u64 *ko, u32 *vo, int p) {
        kern<<<(m + NT * IPT - 1) / (NT * IPT), NT, sm>>>(ki, vi, ko, vo, m, p * 8, b.hist + p * 256, b.status, g_epoch);
    }, NT, IPT, sm, b, m, reps, dist);
}
// the same pass with every tile's prefix taken from the status words an identical untimed launch
// left behind: correct output, no look-back -- the ceiling of the kernel's shape
template <int NT, int IPT, int MINB, int LB>
static void run_config_oracle(const char *name, Bufs &b, u32 m, int reps, int dist)
{
    os_kernel_t warm = k_onesweep_pass<u64, NT, IPT, MINB, LB>;
    os_kernel_t kern = k_onesweep_pass<u64, NT, IPT, MINB, LB, true>;
    const size_t sm = OsSmem<u64, NT, IPT>::bytes;
    CHECK(cudaFuncSetAttribute((const void *)warm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const u32 grid = (m + NT * IPT - 1) / (NT * IPT);
    run_kernel(name, (const void *)kern, [=](const u64 *ki, const u32 *vi, u64 *ko, u32 *vo, int p) {
        kern<<<grid, NT, sm>>>(ki, vi, ko, vo, m, p * 8, b.hist + p * 256, b.status, g_epoch);
    }, NT, IPT, sm, b, m, reps, dist, [=](const u64 *ki, const u32 *vi, u64 *ko, u32 *vo, int p) {
        warm<<<grid, NT, sm>>>(ki, vi, ko, vo, m, p * 8, b.hist + p * 256, b.status, g_epoch);
    });
}

static void run_kernel(const char *name, const void *kern, launcher_t launch, int NT, int IPT, size_t smem_bytes, Bufs &b,
                       u32 m, int reps, int dist, launcher_t pre)
{
    struct { size_t bytes; } Lb = {smem_bytes};
#define L_BYTES Lb.bytes
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_BYTES));
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, L_BYTES));
    cudaFuncAttributes fa;
    CHECK(cudaFuncGetAttributes(&fa, kern));
    const int passes = 4;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    double best[passes], sum[passes];
    for (int p = 0; p < passes; p++) { best[p] = 1e9; sum[p] = 0; }
    u32 errors = 0;
#ifdef OS_PROFILE_PHASES
    unsigned long long zero[32] = {0};
    CHECK(cudaMemcpyToSymbol(g_phase, zero, sizeof zero));
#endif
    for (int r = 0; r < reps + 1; r++) {
        CHECK(cudaMemcpy(b.k[0], b.orig, (size_t)m * 8, cudaMemcpyDeviceToDevice));
        CHECK(cudaMemset(b.hist, 0, (RADIX_MAX_PASSES * RADIX_BINS + RADIX_MAX_PASSES) * 4));
        k_radix_hist<<<148 * 8, 256>>>(b.k[0], m, passes, b.hist);
        k_radix_hist_scan<<<passes, 256>>>(b.hist);
        int cur = 0;
        for (int p = 0; p < passes; p++) {
            g_epoch++;
            if (pre) pre(b.k[cur], p == 0 ? nullptr : b.v[cur], b.k[cur ^ 1], b.v[cur ^ 1], p);  // untimed
            CHECK(cudaEventRecord(e0));
            launch(b.k[cur], p == 0 ? nullptr : b.v[cur], b.k[cur ^ 1], b.v[cur ^ 1], p);
            CHECK(cudaEventRecord(e1));
            CHECK(cudaEventSynchronize(e1));
            CHECK(cudaGetLastError());
            float ms;
            CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0) { sum[p] += ms; if (ms < best[p]) best[p] = ms; }
            if (r == 0 && p == 0) {
                CHECK(cudaMemset(b.errors, 0, 4));
                k_check<<<(m + 255) / 256, 256>>>(b.orig, b.k[1], b.v[1], m, 0, b.errors);
                CHECK(cudaMemcpy(&errors, b.errors, 4, cudaMemcpyDeviceToHost));
            }
            cur ^= 1;
        }
    }
    double tot = 0, totb = 0;
    for (int p = 0; p < passes; p++) { tot += sum[p] / reps; totb += (p == 0 ? 20.0 : 24.0) * m; }
    printf("%-22s dist %d regs %3d smem %6zu occ %d | ms/pass", name, dist, fa.numRegs, (size_t)L_BYTES, occ);
    for (int p = 0; p < passes; p++) printf(" %.3f", sum[p] / reps);
    printf(" | avg %.0f GB/s (%.1f%% of 6552.6) best-pass %.0f GB/s | %s\n", totb / (tot * 1e-3) / 1e9,
           100.0 * totb / (tot * 1e-3) / 1e9 / 6552.6, 24.0 * m / (best[1] * 1e-3) / 1e9, errors ? "WRONG" : "ok");
#ifdef OS_PROFILE_PHASES
    unsigned long long ph[32];
    CHECK(cudaMemcpyFromSymbol(ph, g_phase, sizeof ph));
    const double tiles = (double)((m + NT * IPT - 1) / (NT * IPT)) * passes * (reps + 1);
    static const char *pn[9] = {"ticket+zero", "loads", "rank", "sync(rank)", "digit-scan", "stage", "look-back",
                                "sync(lb)", "write"};
    double all = 0;
    for (int i = 0; i < 9; i++) all += (double)ph[i];
    printf("    cycles/tile:");
    for (int i = 0; i < 9; i++) printf(" %s %.0f", pn[i], ph[i] / tiles);
    printf(" | total %.0f\n", all / tiles);
#endif
    if (errors) printf("    %u order/pairing errors\n", errors);
    CHECK(cudaEventDestroy(e0));
    CHECK(cudaEventDestroy(e1));
}

int main(int argc, char **argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 26;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const u32 m = (1u << lg) - 12345u % (1u << lg);  // not a multiple of any tile size
    Bufs b;
    for (int i = 0; i < 2; i++) {
        CHECK(cudaMalloc(&b.k[i], (size_t)m * 8));
        CHECK(cudaMalloc(&b.v[i], (size_t)m * 4));
    }
    CHECK(cudaMalloc(&b.orig, (size_t)m * 8));
    CHECK(cudaMalloc(&b.hist, (RADIX_MAX_PASSES * RADIX_BINS + RADIX_MAX_PASSES) * 4));
    const size_t tiles = (m + 1023) / 1024;
    CHECK(cudaMalloc(&b.status, tiles * 256 * 8));
    CHECK(cudaMemset(b.status, 0, tiles * 256 * 8));
    CHECK(cudaMalloc(&b.errors, 4));
    printf("m = %u pairs, %d reps\n", m, reps);
    for (int dist = 0; dist < 2; dist++) {
        k_gen<<<(m + 255) / 256, 256>>>(b.orig, m, dist, 42);
        CHECK(cudaDeviceSynchronize());
#define RUN(NT, IPT, MINB, LB) run_config<NT, IPT, MINB, LB>(#NT "x" #IPT " minb" #MINB " lb" #LB, b, m, reps, dist)
#define RUNO(NT, IPT, MINB, LB) run_config_oracle<NT, IPT, MINB, LB>("no-lookback " #NT "x" #IPT " minb" #MINB, b, m, reps, dist)
        RUN(384, 12, 3, 4);
        RUN(384, 12, 3, 2);
        RUNO(384, 12, 3, 4);
        RUNO(512, 8, 3, 4);
        RUNO(384, 8, 4, 4);
        RUNO(256, 16, 3, 4);
#undef RUN
    }
    return 0;
}
