// radix.cuh -- LSD "onesweep" radix sort of (u64 key, u32 value) pairs, 8-bit digits.
//
// One launch per digit.  CTA b owns tile b (tiles are dispatched in index order), ranks its
// keys with warp-level ballot masks into per-warp digit counters, publishes the tile's
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

#define OS_TILE_MAX 4096          // status array is sized for the smallest tile in use (2048)
#define OS_TILE_MIN 2048
#define OS_LOOKBACK 8             // status words fetched per look-back step

#define OS_FLAG_AGG 1ull
#define OS_FLAG_PREFIX 2ull

#define OS_PHASE_INIT() PH_INIT()
#define OS_PHASE(i_) PH(i_)

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
// NT threads x IPT keys per thread = one tile; MINB = CTAs per SM the register budget allows;
// LB = status words fetched per look-back step.
//
// Shape of the kernel, and what the per-phase cycle counters (OS_PROFILE_PHASES,
// tests/bench_onesweep.cu) said about the first version (512 x 8, MATCH.ANY ranking, atomic
// ticket, 29 k cycles per 4096-key tile on uniform digits):
//   * ranking took 7.3 k cycles: MATCH.ANY occupies its unit for ~100 cycles per warp
//     instruction when the 32 digits differ.  Peers are now the AND over the 8 digit bits of
//     (ballot of the bit, complemented where my bit is 0): fixed cost, 4.1 k cycles.
//   * the ticket (an exposed L2 atomic round trip, 1.6 k cycles) is gone: tile = blockIdx.x.
//     CTAs are dispatched in index order -- the assumption cub::DeviceScan's look-back makes
//     as well -- so every tile a look-back waits for is resident or finished.
//   * the look-back (5.5 k cycles) is mostly a wait for the slowest of the ~85 predecessor
//     tiles that have no prefix yet, not for status loads: windows of 16 or 32 words, a
//     cooperative two-group window and a flag-word + 16-bit-aggregate-row protocol all
//     measured slower or equal.  It runs after keys and values have left
//     the registers for shared memory -- that staging needs only tile-local offsets.
//   * three CTAs of 384 threads per SM (56 registers) overlap the phases better than two of 512.
static __device__ __forceinline__ u64 ldg_stream_key(const u64 *p) { return ldg_stream_u64(p); }
static __device__ __forceinline__ u32 ldg_stream_key(const u32 *p) { return ldg_stream_u32(p); }

// K = key type: u64 for the sorts of the doubling, u32 for the binning pass of the rank scatter
template <typename K, int NT, int IPT>
struct OsSmem {
    static constexpr int TILE = NT * IPT, NW = NT / 32;
    static constexpr size_t keys = 0;                                   // K[TILE]
    static constexpr size_t vals = keys + sizeof(K) * TILE;             // u32[TILE]
    static constexpr size_t wcnt = vals + sizeof(u32) * TILE;           // u16[NW][256]
    static constexpr size_t dstart = wcnt + sizeof(u16) * NW * RADIX_BINS;  // u32[256]
    static constexpr size_t adj = dstart + sizeof(u32) * RADIX_BINS;    // u32[256]
    static constexpr size_t wsum = adj + sizeof(u32) * RADIX_BINS;      // u32[8]
    static constexpr size_t bytes = wsum + sizeof(u32) * 8;
};

template <typename K, int NT, int IPT, int MINB, int LB>
__global__ void __launch_bounds__(NT, MINB)
k_onesweep_pass(const K *__restrict__ kin, const u32 *__restrict__ vin, K *__restrict__ kout,
                u32 *__restrict__ vout, u32 m, u32 shift, const u32 *__restrict__ binbase,
                u64 *__restrict__ status, u32 epoch)
{
    using L = OsSmem<K, NT, IPT>;
    constexpr int TILE = L::TILE, NW = L::NW;
    static_assert(NT >= RADIX_BINS && TILE <= 65536, "one thread per digit; 16-bit tile positions");
    extern __shared__ __align__(16) u8 smem[];
    K *s_keys = (K *)(smem + L::keys);
    u32 *s_vals = (u32 *)(smem + L::vals);
    u16(*s_wcnt)[RADIX_BINS] = (u16(*)[RADIX_BINS])(smem + L::wcnt);
    u32 *s_dstart = (u32 *)(smem + L::dstart);
    u32 *s_adj = (u32 *)(smem + L::adj);
    u32 *s_wsum = (u32 *)(smem + L::wsum);

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    OS_PHASE_INIT();
    for (u32 i = tid; i < NW * RADIX_BINS / 2; i += NT) ((u32 *)s_wcnt)[i] = 0;
    __syncthreads();
    OS_PHASE(0);
    const u32 tile = blockIdx.x;
    const u32 base = tile * TILE;
    const u32 cnt = min((u32)TILE, m - base);
    const u32 wbase = base + warp * (32 * IPT);

    // warp-striped loads: slot j of lane l is tile position warp*32*IPT + j*32 + l
    K key[IPT];
    u32 val[IPT];
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 g = wbase + j * 32 + lane;
        key[j] = (g < m) ? ldg_stream_key(kin + g) : (K)~(K)0;  // pads: digit 255, last in tile order
    }
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 g = wbase + j * 32 + lane;
        val[j] = (g < m) ? (vin ? ldg_stream_u32(vin + g) : g) : 0u;
    }
#ifdef BWTS_PROFILE_PHASES
    if (key[IPT - 1] == (K)0x123456789abcdefull && val[IPT - 1] == 0x1234567u) g_phase[15] = 1;  // wait for the loads
#endif
    OS_PHASE(1);

    // rank inside the warp, in slot order (stable)
    u16 *wc = s_wcnt[warp];
    const u32 lt = lanemask_lt();
    u16 rnk[IPT];
    {
        u32 peers[IPT];  // lanes whose digit equals mine
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 dj = (u32)(key[j] >> shift) & (RADIX_BINS - 1);
            u32 p = FULL_MASK;
#pragma unroll
            for (int b = 0; b < RADIX_BITS; b++) {
                const u32 bit = (dj >> b) & 1u;
                const u32 bal = __ballot_sync(FULL_MASK, bit);
                p &= bal ^ (bit - 1u);  // lanes whose bit b equals mine
            }
            peers[j] = p;
        }
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 dj = (u32)(key[j] >> shift) & (RADIX_BINS - 1);
            const int leader = __ffs(peers[j]) - 1;
            u32 before = 0;
            if ((int)lane == leader) {
                before = wc[dj];
                wc[dj] = (u16)(before + __popc(peers[j]));
            }
            before = __shfl_sync(FULL_MASK, before, leader);
            rnk[j] = (u16)(before + __popc(peers[j] & lt));
            __syncwarp();
        }
    }
    OS_PHASE(2);
    __syncthreads();
    OS_PHASE(3);

    // one thread per digit: exclusive offsets of the warps, tile count, publish the aggregate
    u32 blockcnt = 0, dsum = 0;
    const u32 d = tid;
    u64 *my = status + (u64)tile * RADIX_BINS + d;
    if (tid < RADIX_BINS) {
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const u32 c = s_wcnt[w][d];
            s_wcnt[w][d] = (u16)blockcnt;
            blockcnt += c;
        }
        st_relaxed_u64(my, os_pack(epoch, tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG, blockcnt));
        const u32 incl = warp_incl_sum(blockcnt);
        if (lane == 31) s_wsum[warp] = incl;
        dsum = incl - blockcnt;  // exclusive sum inside my warp of digits
    }
    __syncthreads();
    if (tid < RADIX_BINS) {
#pragma unroll
        for (int w = 0; w < RADIX_BINS / 32; w++)
            if (w < (int)warp) dsum += s_wsum[w];
        s_dstart[d] = dsum;  // where the digit's run starts inside the sorted tile
    }
    __syncthreads();
    OS_PHASE(4);

    // keys and values -> shared memory in tile-sorted order (frees their registers)
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 dj = (u32)(key[j] >> shift) & (RADIX_BINS - 1);
        const u32 pos = s_dstart[dj] + wc[dj] + rnk[j];
        s_keys[pos] = key[j];
        s_vals[pos] = val[j];
    }
    OS_PHASE(5);

    // decoupled look-back over the earlier tiles, LB status words per step
    if (tid < RADIX_BINS) {
        u32 excl = 0;
        if (tile != 0) {
            int t = (int)tile - 1;
            bool found = false;
            while (!found) {
                u64 v[LB];
#pragma unroll
                for (int q = 0; q < LB; q++) {
                    const int tt = t - q;
                    v[q] = (tt >= 0) ? ld_relaxed_u64(status + (u64)tt * RADIX_BINS + d)
                                     : os_pack(epoch, OS_FLAG_PREFIX, 0);
                }
                int used = 0;
#pragma unroll
                for (int q = 0; q < LB; q++) {
                    if (found || used != q) continue;          // stopped earlier in this window
                    if ((u32)(v[q] >> 34) != epoch) continue;  // not published yet: re-read from here
                    excl += (u32)v[q];
                    used = q + 1;
                    if (((v[q] >> 32) & 3ull) == OS_FLAG_PREFIX) found = true;
                }
                t -= used;
            }
            st_relaxed_u64(my, os_pack(epoch, OS_FLAG_PREFIX, excl + blockcnt));
        }
        s_adj[d] = __ldg(binbase + d) + excl - dsum;
    }
    OS_PHASE(6);
    __syncthreads();
    OS_PHASE(7);

    // every digit run goes out with consecutive threads on consecutive addresses
#pragma unroll
    for (int q = 0; q < IPT; q++) {
        const u32 s = q * NT + tid;
        if (s < cnt) {
            const K k = s_keys[s];
            const u32 dst = s + s_adj[(u32)(k >> shift) & (RADIX_BINS - 1)];
            kout[dst] = k;
            vout[dst] = s_vals[s];
        }
    }
    OS_PHASE(8);
}

