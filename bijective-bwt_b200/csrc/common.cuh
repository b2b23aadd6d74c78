// common.cuh -- shared device helpers for libbwts_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t  i32;

#define FULL_MASK 0xffffffffu

// Per-phase cycle counters (diagnostic builds only: -DBWTS_PROFILE_PHASES, `make profile-lib`, or
// OS_PROFILE_PHASES in tests/bench_onesweep.cu).  Thread 0 of every CTA adds the cycles since
// its previous mark to g_phase[slot]; slots 0-8 onesweep, 16-27 re-rank.
#if defined(OS_PROFILE_PHASES) && !defined(BWTS_PROFILE_PHASES)
#define BWTS_PROFILE_PHASES
#endif
#ifdef BWTS_PROFILE_PHASES
__device__ unsigned long long g_phase[32];
#define PH_INIT() long long ph_t_prev__ = clock64()
#define PH(i_)                                                                \
    do {                                                                      \
        if (threadIdx.x == 0) {                                               \
            const long long t__ = clock64();                                  \
            atomicAdd(&g_phase[i_], (unsigned long long)(t__ - ph_t_prev__)); \
            ph_t_prev__ = t__;                                                \
        }                                                                     \
    } while (0)
#else
#define PH_INIT() do { } while (0)
#define PH(i_) do { } while (0)
#endif
#define NONE32 0xffffffffu

static __device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
static __device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// single-word status words of the decoupled look-back scans: relaxed gpu-scope accesses
static __device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming (read-once) loads: keep them out of L1
static __device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
static __device__ __forceinline__ u64 ldg_stream_u64(const u64 *p)
{
    u64 r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
static __device__ __forceinline__ u32 ldg_stream_u32(const u32 *p)
{
    u32 r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// 8 bytes at an arbitrary address, assembled from aligned words.  The caller guarantees
// that p+16 does not run past the end of the buffer.
static __device__ __forceinline__ u64 load8_unaligned(const u8 *p)
{
    const uintptr_t a = (uintptr_t)p;
    const u64 *w = (const u64 *)(a & ~(uintptr_t)7);
    const u32 sh = (u32)(a & 7) * 8;
    u64 lo = w[0];
    if (sh == 0) return lo;
    u64 hi = w[1];
    return (lo >> sh) | (hi << (64 - sh));
}

// inclusive warp scan (sum)
static __device__ __forceinline__ u32 warp_incl_sum(u32 v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(FULL_MASK, v, o);
        if (lane_id() >= (u32)o) v += y;
    }
    return v;
}
static __device__ __forceinline__ u32 warp_incl_max(u32 v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(FULL_MASK, v, o);
        if (lane_id() >= (u32)o) v = max(v, y);
    }
    return v;
}
static __device__ __forceinline__ u32 warp_sum(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
static __device__ __forceinline__ u32 warp_max(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
static __device__ __forceinline__ u32 warp_min(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// Lyndon factor table: FS[0..F] ascending factor starts with FS[F] = n, plus a coarse
// index cidx[b] = index of the factor containing position b << CB (cidx has nblk+1
// entries, the last one = F-1).  Returns f with FS[f] <= i < FS[f+1].
#define COARSE_BITS 12
static __device__ __forceinline__ u32 factor_of(const u32 *__restrict__ FS, const u32 *__restrict__ cidx, u32 i)
{
    u32 b = i >> COARSE_BITS;
    u32 lo = __ldg(cidx + b), hi = __ldg(cidx + b + 1);
    while (lo < hi) {
        u32 mid = (lo + hi + 1) >> 1;
        if (__ldg(FS + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}
