// radix.cuh -- LSD "onesweep" radix sort of (u64 key, u32 value) pairs, 8-bit digits.
//
// One launch per digit.  CTA b owns tile b (tiles are dispatched in index order): the raw tile
// is bulk-copied into shared memory, the keys are ranked by their digit byte with warp-level
// ballot masks into per-warp digit counters, the tile's 256 digit counts are published, the
// global offsets come from a decoupled look-back over the earlier tiles' status words (single
// 64-bit words carrying epoch | flag | count, so no reset between passes), and every digit run
// is written with consecutive threads on consecutive addresses.
//
// The digit histograms of all passes are taken up front in one sweep (k_radix_hist).
#pragma once
#include "common.cuh"

#define RADIX_BITS 8
#define RADIX_BINS 256
#define RADIX_MAX_PASSES 8

#define OS_TILE_MIN 4096          // the status array is sized for the smallest tile a configuration may use

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
// K = key type: u64 for the sorts of the doubling, u32 for the binning passes of the rank
// scatter and of the large-output emit.  NT threads x IPT keys per thread = one tile; MINB = CTAs
// per SM; LB = status words fetched per look-back step.
//
// Shape of the kernel (per-phase cycle counters: -DOS_PROFILE_PHASES, tests/bench_onesweep.cu;
// the history of the numbers is in DESIGN.md sections 4.4 / 4.5):
//   * the raw tile lands in shared memory through two bulk copies (cp.async.bulk = the TMA unit,
//     completion counted on an mbarrier) issued by one thread.  Keys and values never occupy
//     registers: the first version held 8 of each per thread from the load to the staging
//     (56 registers, 3 CTAs of 3072 keys per SM, 44 % of the HBM peak); this one runs 3 CTAs of
//     4608 keys at the same register count and 54-56 %.
//   * ranking reads only the digit byte of each key.  Peers of a key = AND over the 8 digit bits
//     of (ballot of the bit, complemented where my bit is 0): MATCH.ANY held its unit ~100 cycles
//     per warp instruction when the 32 digits differ (7.3 k of 29 k cycles per tile).
//   * tile = blockIdx.x.  CTAs are dispatched in index order -- the assumption cub::DeviceScan's
//     look-back makes as well -- so every tile a look-back waits for is resident or finished;
//     an atomic ticket was an exposed L2 round trip per tile.
//   * the tile-sorted order is a 16-bit permutation (sorted slot -> raw position); the write
//     phase gathers keys and values through it, consecutive threads on consecutive addresses
//     of every digit run.
//   * decoupled look-back with single 64-bit status words (epoch | flag | count, no reset between
//     passes).  Bigger tiles mean fewer tiles in flight and a shorter walk; windows of 2-4 words beat
//     8 and 16 (a wider window mostly fetches words behind the tile that ends the walk).  With every
//     tile's prefix handed to it (ORACLE, timing only) the pass runs 10 % faster: that is all the
//     look-back still costs.
template <typename K, int NT, int IPT>
struct OsSmem {
    static constexpr int TILE = NT * IPT, NW = NT / 32;
    static constexpr size_t keys = 0;                                    // K[TILE], raw order
    static constexpr size_t vals = keys + sizeof(K) * TILE;              // u32[TILE], raw order
    static constexpr size_t inv = vals + sizeof(u32) * TILE;             // u16[TILE]: sorted slot -> raw position
    static constexpr size_t wcnt = inv + sizeof(u16) * TILE;             // u16[NW][256]
    static constexpr size_t dstart = wcnt + sizeof(u16) * NW * RADIX_BINS;   // u32[256]
    static constexpr size_t adj = dstart + sizeof(u32) * RADIX_BINS;     // u32[256]
    static constexpr size_t wsum = adj + sizeof(u32) * RADIX_BINS;       // u32[8]
    static constexpr size_t mbar = wsum + sizeof(u32) * 8;               // u64
    static constexpr size_t bytes = mbar + 16;
};

static __device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// MODE 0: the sort pass.  MODE 1 (ORACLE): timing experiment, see below.  MODE 2 (PACK, u32 keys): the
// binning pass of the large-output emit -- `vin` is the TEXT (bytes), the value of element i is
// T[i - 1], and the only output is one packed word per element, (key & low bits) | byte << shift: inside
// a rank bin the high bits of the rank are the bin number, which the position in the stream gives back
// (k_scatter_packed).  14 bytes per element instead of 30 for the (rank, byte) pairs as two u32 streams.
template <typename K, int NT, int IPT, int MINB, int LB, int MODE = 0>
__global__ void __launch_bounds__(NT, MINB)
k_onesweep_pass(const K *__restrict__ kin, const u32 *__restrict__ vin_or_text, K *__restrict__ kout,
                  u32 *__restrict__ vout, u32 m, u32 shift, const u32 *__restrict__ binbase,
                  u64 *__restrict__ status, u32 epoch)
{
    constexpr bool ORACLE = MODE == 1, PACK = MODE == 2;
    static_assert(!PACK || sizeof(K) == 4, "the packed emit pass bins 32-bit ranks");
    const u32 *__restrict__ vin = PACK ? (const u32 *)nullptr : vin_or_text;
    const u8 *__restrict__ text = (const u8 *)vin_or_text;
    using L = OsSmem<K, NT, IPT>;
    constexpr int TILE = L::TILE, NW = L::NW;
    static_assert(NT >= RADIX_BINS && TILE <= 65536 && IPT % 4 == 0, "one thread per digit; 16-bit tile positions");
    extern __shared__ __align__(128) u8 smem[];
    K *s_keys = (K *)(smem + L::keys);
    u32 *s_vals = (u32 *)(smem + L::vals);
    u16 *s_inv = (u16 *)(smem + L::inv);
    u16(*s_wcnt)[RADIX_BINS] = (u16(*)[RADIX_BINS])(smem + L::wcnt);
    u32 *s_dstart = (u32 *)(smem + L::dstart);
    u32 *s_adj = (u32 *)(smem + L::adj);
    u32 *s_wsum = (u32 *)(smem + L::wsum);
    const u32 mbar = smem_addr(smem + L::mbar);

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    OS_PHASE_INIT();
    const u32 tile = blockIdx.x;
    const u32 base = tile * TILE;
    const u32 cnt = min((u32)TILE, m - base);
    const bool bulk = cnt == (u32)TILE;

    if (tid == 0 && bulk) {
        const u32 kbytes = (u32)(sizeof(K) * TILE), vbytes = vin ? (u32)(sizeof(u32) * TILE) : 0u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(kbytes + vbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(s_keys)), "l"(kin + base), "r"(kbytes), "r"(mbar) : "memory");
        if (vin)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(s_vals)), "l"(vin + base), "r"(vbytes), "r"(mbar) : "memory");
    }
    for (u32 i = tid; i < NW * RADIX_BINS / 2; i += NT) ((u32 *)s_wcnt)[i] = 0;
    if (!bulk) {  // last tile: plain loads, pads = all ones (digit 255, last in tile order)
        for (u32 p = tid; p < (u32)TILE; p += NT) {
            s_keys[p] = (p < cnt) ? kin[base + p] : (K)~(K)0;
            if (vin) s_vals[p] = (p < cnt) ? vin[base + p] : 0u;
        }
    }
    if (PACK) {
        // the tile's text bytes, staged in the (otherwise unused) value area: s_txt[b] = T[base - 4 + b], so the byte
        // before element p is s_txt[p + 3].  Whole words while they lie inside the text (base is a multiple of 4, the
        // text pointer 16-byte aligned); the last tile goes byte by byte.
        u8 *s_txt = (u8 *)s_vals;
        if (bulk) {
            const u32 *tw = (const u32 *)text + base / 4;
            for (u32 w = tid; w <= (u32)TILE / 4; w += NT) s_vals[w] = (w == 0 && base == 0) ? 0u : ldg_stream_u32(tw + w - 1);
        } else {
            for (u32 p = tid; p < cnt; p += NT) s_txt[p + 3] = (base + p) ? text[base + p - 1] : (u8)0;
        }
    }
    __syncthreads();  // counters zeroed, barrier initialised (or the plain loads done)
    OS_PHASE(0);
    if (bulk) {
        u32 done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar) : "memory");
    }
    OS_PHASE(1);

    // rank inside the warp, in slot order (stable): only the digit byte of each key is read
    u16 *wc = s_wcnt[warp];
    const u32 lt = lanemask_lt();
    const u32 wpos = warp * (32 * IPT) + lane;  // raw position of my slot 0; slot j is 32 further each
    const u8 *dbytes = (const u8 *)s_keys + (shift >> 3);
    const bool bytewise = (shift & 7u) == 0;
    u32 dpk[IPT / 4];  // my digits, 4 per word
    u16 rnk[IPT];
    {
        u32 peers[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            // byte-aligned digits (all LSD passes) cost one byte load; the binning passes use an
            // arbitrary shift and read the whole key
            const u32 dj = bytewise ? (u32)dbytes[(size_t)(wpos + j * 32) * sizeof(K)]
                                    : ((u32)(s_keys[wpos + j * 32] >> shift) & (RADIX_BINS - 1));
            if ((j & 3) == 0) dpk[j >> 2] = 0;
            dpk[j >> 2] |= dj << (8 * (j & 3));
            u32 p = FULL_MASK;
#pragma unroll
            for (int b = 0; b < RADIX_BITS; b++) {
                const u32 bit = (dj >> b) & 1u;
                const u32 bal = __ballot_sync(FULL_MASK, bit);
                p &= bal ^ (bit - 1u);
            }
            peers[j] = p;
        }
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 dj = (dpk[j >> 2] >> (8 * (j & 3))) & 255u;
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
        if (!ORACLE) st_relaxed_u64(my, os_pack(epoch, tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG, blockcnt));
        const u32 incl = warp_incl_sum(blockcnt);
        if (lane == 31) s_wsum[warp] = incl;
        dsum = incl - blockcnt;
    }
    __syncthreads();
    if (tid < RADIX_BINS) {
#pragma unroll
        for (int w = 0; w < RADIX_BINS / 32; w++)
            if (w < (int)warp) dsum += s_wsum[w];
        s_dstart[d] = dsum;
    }
    __syncthreads();
    OS_PHASE(4);

    // the tile-sorted order as a permutation: sorted slot -> raw position
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 dj = (dpk[j >> 2] >> (8 * (j & 3))) & 255u;
        s_inv[s_dstart[dj] + wc[dj] + rnk[j]] = (u16)(wpos + j * 32);
    }
    OS_PHASE(5);

    if (tid < RADIX_BINS) {
        u32 excl = 0;
        if (ORACLE) {
            // timing experiment (tests/bench_onesweep.cu): the prefix this tile published in an identical
            // earlier launch is still in its status word -- the kernel without any look-back
            excl = (u32)ld_relaxed_u64(my) - blockcnt;
        } else if (tile != 0) {
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
                    if (found || used != q) continue;
                    if ((u32)(v[q] >> 34) != epoch) continue;
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

#pragma unroll
    for (int q = 0; q < IPT; q++) {
        const u32 s = q * NT + tid;
        if (s < cnt) {
            const u32 p = s_inv[s];
            const K k = s_keys[p];
            const u32 dst = s + s_adj[(u32)(k >> shift) & (RADIX_BINS - 1)];
            if (PACK) {
                // element index base + p = text position; position 0 starts a factor (k_emit_heads)
                const u32 byte = (u32)((const u8 *)s_vals)[p + 3];
                vout[dst] = ((u32)k & ((1u << shift) - 1u)) | (byte << shift);
            } else {
                kout[dst] = k;
                vout[dst] = vin ? s_vals[p] : base + p;
            }
        }
    }
    OS_PHASE(8);
}
