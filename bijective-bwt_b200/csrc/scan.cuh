// scan.cuh -- exclusive prefix sums of u32 arrays (reduce / scan-of-tiles / apply).
#pragma once
#include "common.cuh"

// One block of 1024 threads scans `count` values (any count), 4 per thread per sweep.
// in and out may alias.  *total (optional) receives the grand total.
__global__ void __launch_bounds__(1024) k_scan_excl_u32_block(const u32 *in, u32 *out, u32 count, u32 *total)
{
    __shared__ u32 ws[33];
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    u32 carry = 0;  // identical in every thread
    for (u32 base = 0; base < count; base += 4096) {
        const u32 i0 = base + threadIdx.x * 4;
        u32 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = (i0 + q < count) ? in[i0 + q] : 0;
        const u32 mine = v[0] + v[1] + v[2] + v[3];
        const u32 incl = warp_incl_sum(mine);
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const u32 w = ws[lane];
            const u32 wi = warp_incl_sum(w);
            ws[lane] = wi - w;
            if (lane == 31) ws[32] = wi;
        }
        __syncthreads();
        u32 run = carry + ws[warp] + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (i0 + q < count) out[i0 + q] = run;
            run += v[q];
        }
        carry += ws[32];
        __syncthreads();
    }
    if (total && threadIdx.x == 0) *total = carry;
}

// ---- large arrays: tile sums -> (block scan of the sums) -> apply -------------------------
#define SC_TILE 4096  // 256 threads x 16 values
// nonzero != nullptr: the number of non-zero inputs is added to *nonzero (one atomic per tile)
__global__ void __launch_bounds__(256) k_tile_sum_u32(const u32 *__restrict__ in, u32 n, u32 *__restrict__ tile_sum,
                                                      u32 *__restrict__ nonzero)
{
    __shared__ u32 ws[8], wz[8];
    const u32 base = blockIdx.x * SC_TILE;
    u32 s = 0, z = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const u32 i = base + q * 256 + threadIdx.x;
        if (i < n) {
            const u32 v = ldg_stream_u32(in + i);
            s += v;
            z += v != 0;
        }
    }
    s = warp_sum(s);
    z = warp_sum(z);
    if (lane_id() == 0) { ws[threadIdx.x >> 5] = s; wz[threadIdx.x >> 5] = z; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0, tz = 0;
        for (int w = 0; w < 8; w++) { t += ws[w]; tz += wz[w]; }
        tile_sum[blockIdx.x] = t;
        if (nonzero && tz) atomicAdd(nonzero, tz);
    }
}

// out[i] = tile_off[tile] + exclusive sum inside the tile.  Thread t owns 16 consecutive values.
__global__ void __launch_bounds__(256) k_tile_scan_apply_u32(const u32 *__restrict__ in, u32 *__restrict__ out, u32 n,
                                                             const u32 *__restrict__ tile_off)
{
    __shared__ u32 ws[8];
    const u32 base = blockIdx.x * SC_TILE + threadIdx.x * 16;
    u32 v[16];
    u32 s = 0;
    if (base + 16 <= n) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 x = ldg_stream_u4((const uint4 *)(in + base) + q);
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = (base + q < n) ? in[base + q] : 0;
    }
#pragma unroll
    for (int q = 0; q < 16; q++) s += v[q];
    const u32 incl = warp_incl_sum(s);
    if (lane_id() == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 run = tile_off[blockIdx.x] + incl - s;
    for (u32 w = 0; w < (threadIdx.x >> 5); w++) run += ws[w];
    if (base + 16 <= n) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint4 y;
            y.x = run; run += v[4 * q];
            y.y = run; run += v[4 * q + 1];
            y.z = run; run += v[4 * q + 2];
            y.w = run; run += v[4 * q + 3];
            ((uint4 *)(out + base))[q] = y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            if (base + q < n) out[base + q] = run;
            run += v[q];
        }
    }
}
