// radix.cuh -- LSD "onesweep" radix sort of (u64 key, u32 value) pairs, 8-bit digits.
//
// One launch per digit.  Each CTA takes the next tile from an atomic ticket, ranks its
// keys with warp-level match masks into per-warp digit counters, publishes the tile's
// 256 digit counts, resolves its global offsets by decoupled look-back over the earlier
// tiles' status words (single 64-bit words carrying epoch | flag | count, so no reset
// between passes), stages keys and values through shared memory in tile-sorted order
// and writes every digit run with consecutive threads on consecutive addresses.
//
// The digit histograms of all passes are taken up front in one sweep (k_radix_hist).
#pragma once
#include "common.cuh"

#define RADIX_BITS 8
#define RADIX_BINS 256
#define RADIX_MAX_PASSES 8

#define OS_NT 256                 // threads per CTA
#define OS_IPT 16                 // keys per thread
#define OS_TILE (OS_NT * OS_IPT)  // 4096 keys per tile
#define OS_NW (OS_NT / 32)

#define OS_FLAG_AGG 1ull
#define OS_FLAG_PREFIX 2ull

static __device__ __forceinline__ u64 os_pack(u32 epoch, u64 flag, u32 value)
{
    return ((u64)epoch << 34) | (flag << 32) | (u64)value;
}

// ---- histograms of every digit, one sweep over the keys ---------------------------------
// ghist: [RADIX_MAX_PASSES][256], zeroed by the host before the launch.
__global__ void __launch_bounds__(256) k_radix_hist(const u64 *__restrict__ keys, u32 m, int passes,
                                                    u32 *__restrict__ ghist)
{
    __shared__ u32 sh[RADIX_MAX_PASSES][RADIX_BINS];
    for (u32 i = threadIdx.x; i < RADIX_MAX_PASSES * RADIX_BINS; i += blockDim.x) ((u32 *)sh)[i] = 0;
    __syncthreads();
    const u32 lane = lane_id();
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 gb = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); gb < m; gb += stride) {
        const u32 g = gb + lane;
        const bool valid = g < m;
        const u64 k = valid ? ldg_stream_u64(keys + g) : 0ull;
        const bool whole = __all_sync(FULL_MASK, valid);
        for (int p = 0; p < passes; p++) {
            const u32 d = (u32)(k >> (p * RADIX_BITS)) & (RADIX_BINS - 1);
            const u32 d0 = __shfl_sync(FULL_MASK, d, 0);
            if (whole && __all_sync(FULL_MASK, d == d0)) {
                if (lane == 0) atomicAdd(&sh[p][d0], 32u);
            } else if (valid) {
                atomicAdd(&sh[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < (u32)passes * RADIX_BINS; i += blockDim.x) {
        u32 c = ((u32 *)sh)[i];
        if (c) atomicAdd(ghist + i, c);
    }
}

// exclusive scan of each pass's 256 counts, in place (grid = passes, block = 256)
__global__ void __launch_bounds__(256) k_radix_hist_scan(u32 *__restrict__ ghist)
{
    __shared__ u32 wsum[8];
    u32 *h = ghist + blockIdx.x * RADIX_BINS;
    const u32 c = h[threadIdx.x];
    const u32 incl = warp_incl_sum(c);
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 off = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); w++) off += wsum[w];
    h[threadIdx.x] = off + incl - c;
}

// ---- one onesweep pass -----------------------------------------------------------------
// vin == nullptr means "values are the element indices" (first pass of the initial sort).
__global__ void __launch_bounds__(OS_NT) k_onesweep_pass(const u64 *__restrict__ kin, const u32 *__restrict__ vin,
                                                         u64 *__restrict__ kout, u32 *__restrict__ vout, u32 m,
                                                         u32 shift, const u32 *__restrict__ binbase,
                                                         u64 *__restrict__ status, u32 *__restrict__ ticket,
                                                         u32 epoch)
{
    __shared__ __align__(16) u64 s_keys[OS_TILE];  // reused for the values
    __shared__ u32 s_wcnt[OS_NW][RADIX_BINS];
    __shared__ u32 s_dstart[RADIX_BINS];
    __shared__ u32 s_adj[RADIX_BINS];
    __shared__ u8 s_dig[OS_TILE];
    __shared__ u32 s_wsum[OS_NW];
    __shared__ u32 s_tile;

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (u32 i = tid; i < OS_NW * RADIX_BINS; i += OS_NT) ((u32 *)s_wcnt)[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * OS_TILE;
    const u32 cnt = min((u32)OS_TILE, m - base);
    const u32 wbase = base + warp * (32 * OS_IPT);

    // warp-striped load: slot j of lane l is tile position warp*512 + j*32 + l
    u64 key[OS_IPT];
#pragma unroll
    for (int j = 0; j < OS_IPT; j++) {
        const u32 g = wbase + j * 32 + lane;
        key[j] = (g < m) ? ldg_stream_u64(kin + g) : ~0ull;  // pads: digit 255, last in tile order
    }

    // rank inside the warp, in slot order (stable)
    u32 *wc = s_wcnt[warp];
    const u32 lt = lanemask_lt();
    u16 rnk[OS_IPT];
#pragma unroll
    for (int j = 0; j < OS_IPT; j++) {
        const u32 d = (u32)(key[j] >> shift) & (RADIX_BINS - 1);
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const int leader = __ffs(peers) - 1;
        u32 before = 0;
        if ((int)lane == leader) {
            before = wc[d];
            wc[d] = before + __popc(peers);
        }
        before = __shfl_sync(FULL_MASK, before, leader);
        rnk[j] = (u16)(before + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // per digit (thread d): exclusive offsets of the warps, tile count
    const u32 d = tid;
    u32 blockcnt = 0;
#pragma unroll
    for (int w = 0; w < OS_NW; w++) {
        const u32 c = s_wcnt[w][d];
        s_wcnt[w][d] = blockcnt;
        blockcnt += c;
    }

    // publish, then look back over earlier tiles
    u64 *my = status + (u64)tile * RADIX_BINS + d;
    u32 excl = 0;
    if (tile == 0) {
        st_relaxed_u64(my, os_pack(epoch, OS_FLAG_PREFIX, blockcnt));
    } else {
        st_relaxed_u64(my, os_pack(epoch, OS_FLAG_AGG, blockcnt));
        const u64 *p = my - RADIX_BINS;
        for (;;) {
            const u64 v = ld_relaxed_u64(p);
            if ((u32)(v >> 34) != epoch) continue;  // not written in this pass yet
            excl += (u32)v;
            if (((v >> 32) & 3ull) == OS_FLAG_PREFIX) break;
            p -= RADIX_BINS;
        }
        st_relaxed_u64(my, os_pack(epoch, OS_FLAG_PREFIX, excl + blockcnt));
    }

    // where each digit's run starts inside the sorted tile
    const u32 incl = warp_incl_sum(blockcnt);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < OS_NW; w++)
        if (w < (int)warp) woff += s_wsum[w];
    const u32 dstart = woff + incl - blockcnt;
    s_dstart[d] = dstart;
    s_adj[d] = __ldg(binbase + d) + excl - dstart;
    __syncthreads();

    // keys -> shared memory in tile-sorted order
#pragma unroll
    for (int j = 0; j < OS_IPT; j++) {
        const u32 dj = (u32)(key[j] >> shift) & (RADIX_BINS - 1);
        const u32 pos = s_dstart[dj] + wc[dj] + rnk[j];
        s_keys[pos] = key[j];
        s_dig[pos] = (u8)dj;
        rnk[j] = (u16)pos;
    }
    __syncthreads();
#pragma unroll 4
    for (u32 s = tid; s < cnt; s += OS_NT) kout[s + s_adj[s_dig[s]]] = s_keys[s];
    __syncthreads();

    // values take the same route
    u32 *s_vals = (u32 *)s_keys;
#pragma unroll
    for (int j = 0; j < OS_IPT; j++) {
        const u32 g = wbase + j * 32 + lane;
        if (g < m) s_vals[rnk[j]] = vin ? ldg_stream_u32(vin + g) : g;
    }
    __syncthreads();
#pragma unroll 4
    for (u32 s = tid; s < cnt; s += OS_NT) vout[s + s_adj[s_dig[s]]] = s_vals[s];
}
