// forward.cuh -- forward BWTS kernels around the radix sort: alphabet compaction, packed
// initial keys, one prefix-doubling round (key build, re-rank + live-set compaction), emit.
//
// State of the doubling (all u32, live arrays indexed by position j in the live array):
//   rank[i]   for every text position: number of rotations strictly smaller at the current
//             depth (= global slot of the head of i's group)
//   idx[j]    text position of the j-th live rotation; live = its group has > 1 member;
//             the live array is ordered by rank, groups are contiguous
//   grp[j]    rank of the group that owns live slot j      (position-indexed, not sorted)
//   gst[j]    live-array offset at which that group starts (position-indexed, not sorted)
// A round sorts (gst | rank[succ^k(idx)]) inside the live array, splits groups where the
// sorted keys change, drops rotations that became unique, and doubles k.
#pragma once
#include "common.cuh"

// ---- alphabet compaction ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_byte_presence(const u8 *__restrict__ T, u32 n, u32 *__restrict__ present)
{
    __shared__ u8 seen[256];  // racing stores of the same value are fine
    seen[threadIdx.x] = 0;
    __syncthreads();
    const u32 nvec = n / 16;
    for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
        const uint4 x = ldg_stream_u4((const uint4 *)T + v);
        const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int s = 0; s < 32; s += 8) seen[(w[q] >> s) & 255] = 1;
    }
    if (blockIdx.x == 0)
        for (u32 p = nvec * 16 + threadIdx.x; p < n; p += blockDim.x) seen[T[p]] = 1;
    __syncthreads();
    const u32 bits = __ballot_sync(FULL_MASK, seen[threadIdx.x] != 0);  // warp w covers bytes 32w..32w+31
    if (lane_id() == 0 && bits) atomicOr(present + (threadIdx.x >> 5), bits);
}

// code[c] = number of present bytes below c; *sigma = number of present bytes
__global__ void k_code_table(const u32 *__restrict__ present, u8 *__restrict__ code, u32 *__restrict__ sigma)
{
    const u32 c = threadIdx.x;  // 256 threads
    u32 below = 0;
    for (u32 w = 0; w < (c >> 5); w++) below += __popc(present[w]);
    below += __popc(present[c >> 5] & ((1u << (c & 31)) - 1));
    code[c] = (u8)below;
    if (c == 255) *sigma = below + ((present[7] >> 31) & 1);
}

// ---- initial keys: k0 packed symbols of the rotation starting at i -------------------------
// Thread owns 8 consecutive positions; inside one factor and away from its end the window
// slides by one symbol per position.
__global__ void __launch_bounds__(256) k_init_keys(const u8 *__restrict__ T, u32 n, const u32 *__restrict__ FS,
                                                   const u32 *__restrict__ cidx, const u8 *__restrict__ code,
                                                   u32 bits, u32 k0, u64 *__restrict__ keys)
{
    __shared__ u8 s_code[256];
    s_code[threadIdx.x] = code[threadIdx.x];
    __syncthreads();
    const u32 i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= n) return;
    const u32 iend = min(n, i0 + 8);
    const u64 mask = (k0 * bits >= 64) ? ~0ull : ((1ull << (k0 * bits)) - 1);
    u32 i = i0;
    while (i < iend) {
        const u32 f = factor_of(FS, cidx, i);
        const u32 s = FS[f], e = FS[f + 1];
        // first key of this stretch, symbol by symbol with cyclic wrap
        u64 key = 0;
        u32 pos = i;
        for (u32 c = 0; c < k0; c++) {
            key = (key << bits) | s_code[T[pos]];
            pos = (pos + 1 == e) ? s : pos + 1;
        }
        keys[i] = key;
        i++;
        // slide while the window [i, i+k0) stays inside the factor
        while (i < iend && i < e && (u64)i + k0 <= e) {
            key = ((key << bits) | s_code[T[i + k0 - 1]]) & mask;
            keys[i] = key;
            i++;
        }
    }
}

// suffix-array variant: no factors, symbols are code+1, positions past the end read as 0
__global__ void __launch_bounds__(256) k_init_keys_linear(const u8 *__restrict__ T, u32 n, const u8 *__restrict__ code,
                                                          u32 bits, u32 k0, u64 *__restrict__ keys)
{
    __shared__ u8 s_code[256];
    s_code[threadIdx.x] = code[threadIdx.x];
    __syncthreads();
    const u32 i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= n) return;
    const u32 iend = min(n, i0 + 8);
    const u64 mask = (k0 * bits >= 64) ? ~0ull : ((1ull << (k0 * bits)) - 1);
    u64 key = 0;
    for (u32 c = 0; c < k0; c++) {
        const u64 p = (u64)i0 + c;
        key = (key << bits) | (p < n ? (u64)s_code[T[p]] + 1 : 0ull);
    }
    keys[i0] = key;
    for (u32 i = i0 + 1; i < iend; i++) {
        const u64 p = (u64)i + k0 - 1;
        key = ((key << bits) | (p < n ? (u64)s_code[T[p]] + 1 : 0ull)) & mask;
        keys[i] = key;
    }
}

// ---- key build of one doubling round ---------------------------------------------------------
// key[j] = gst[j] << kb | rank[succ^k(idx[j])]
__global__ void __launch_bounds__(256) k_build_keys(const u32 *__restrict__ idx, const u32 *__restrict__ gst, u32 m,
                                                    const u32 *__restrict__ rank, const u32 *__restrict__ FS,
                                                    const u32 *__restrict__ cidx, u32 k, u32 kb,
                                                    u64 *__restrict__ keys)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const u32 i = ldg_stream_u32(idx + j);
    const u32 f = factor_of(FS, cidx, i);
    const u32 s = __ldg(FS + f), len = __ldg(FS + f + 1) - s;
    u32 o = i - s;
    if (len > 1) {
        const u32 kk = (k < len) ? k : k % len;
        o += kk;                      // < 2 * len <= 2^31
        if (o >= len) o -= len;
    }
    const u32 r = __ldg(rank + s + o);
    keys[j] = ((u64)ldg_stream_u32(gst + j) << kb) | (u64)r;
}

// linear (suffix-array) variant: succ^k(i) = i + k, the end of the text is the smallest symbol
__global__ void __launch_bounds__(256) k_build_keys_linear(const u32 *__restrict__ idx, const u32 *__restrict__ gst,
                                                           u32 m, const u32 *__restrict__ rank, u32 n, u32 k, u32 kb,
                                                           u64 *__restrict__ keys)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const u32 i = ldg_stream_u32(idx + j);
    const u64 t = (u64)i + k;
    const u32 r = (t < n) ? __ldg(rank + (u32)t) + 1 : 0;
    keys[j] = ((u64)ldg_stream_u32(gst + j) << kb) | (u64)r;
}

// ---- re-rank + compaction -----------------------------------------------------------------------
#define RR_NT 256
#define RR_IPT 8
#define RR_TILE (RR_NT * RR_IPT)
#define RR_FLAG_AGG 1ull
#define RR_FLAG_PREFIX 2ull
// status word: flag[63:62] | (last head position + 1)[61:31] | kept count[30:0]
static __device__ __forceinline__ u64 rr_pack(u64 flag, u32 headp1, u32 keep)
{
    return (flag << 62) | ((u64)headp1 << 31) | (u64)keep;
}

struct RerankCounters {  // zeroed before each launch
    u32 ticket;
    u32 heads;   // groups after the split (all of them)
    u32 kheads;  // groups that stay live
    u32 kept;    // live elements after compaction
};

// finalize != 0: every element becomes its own group (used once ties are known to be final).
__global__ void __launch_bounds__(RR_NT) k_rerank(const u64 *__restrict__ keys, const u32 *__restrict__ idx,
                                                  const u32 *__restrict__ grp, const u32 *__restrict__ gst, u32 m,
                                                  int finalize, u32 *__restrict__ rank, u32 *__restrict__ idx_out,
                                                  u32 *__restrict__ grp_out, u32 *__restrict__ gst_out,
                                                  u64 *__restrict__ status, RerankCounters *__restrict__ ctr)
{
    __shared__ u8 s_head[RR_TILE + 1];
    __shared__ u32 s_wh[RR_NT / 32], s_wk[RR_NT / 32];
    __shared__ u32 s_tile, s_exh, s_exk;

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ctr->ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile, base = tile * RR_TILE;
    const u32 cnt = min((u32)RR_TILE, m - base);

    // head flags for positions base .. base+cnt (one past the tile), coalesced
    for (u32 s = tid; s <= cnt; s += RR_NT) {
        const u32 j = base + s;
        u8 h = 1;
        if (j < m && !finalize) {
            h = (__ldg(gst + j) == j);
            if (!h) h = ldg_stream_u64(keys + j) != __ldg(keys + j - 1);
        }
        s_head[s] = h;
    }
    __syncthreads();

    // thread owns RR_IPT consecutive slots
    const u32 s0 = tid * RR_IPT;
    u32 lasth = 0, nkeep = 0, nhead = 0, nkhead = 0;  // lasth = position+1 of the last head in my slots
    u32 hbits = 0, kbits = 0;
#pragma unroll
    for (int q = 0; q < RR_IPT; q++) {
        const u32 s = s0 + q;
        if (s < cnt) {
            const u32 h = s_head[s], hn = s_head[s + 1];
            const u32 keep = !(h && hn);
            hbits |= h << q;
            kbits |= keep << q;
            if (h) { lasth = base + s + 1; nhead++; nkhead += keep; }
            nkeep += keep;
        }
    }
    // block-wide inclusive scans of (max lasth, sum nkeep)
    const u32 ih = warp_incl_max(lasth), ik = warp_incl_sum(nkeep);
    if (lane == 31) { s_wh[warp] = ih; s_wk[warp] = ik; }
    const u32 th = warp_sum(nhead), tkh = warp_sum(nkhead);
    if (lane == 0 && (th | tkh)) { atomicAdd(&ctr->heads, th); atomicAdd(&ctr->kheads, tkh); }
    __syncthreads();
    u32 offh = 0, offk = 0, toth = 0, totk = 0;
#pragma unroll
    for (int w = 0; w < RR_NT / 32; w++) {
        if (w < (int)warp) { offh = max(offh, s_wh[w]); offk += s_wk[w]; }
        toth = max(toth, s_wh[w]);
        totk += s_wk[w];
    }

    // decoupled look-back on (max, sum), by warp 0
    if (warp == 0) {
        u32 exh = 0, exk = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(status, rr_pack(RR_FLAG_PREFIX, toth, totk));
        } else {
            if (lane == 0) st_relaxed_u64(status + tile, rr_pack(RR_FLAG_AGG, toth, totk));
            int t = (int)tile - 1;
            for (;;) {
                const int q = t - (int)lane;
                const u64 v = (q >= 0) ? ld_relaxed_u64(status + q) : rr_pack(RR_FLAG_PREFIX, 0, 0);
                const u32 flag = (u32)(v >> 62);
                const u32 empties = __ballot_sync(FULL_MASK, flag == 0);
                const u32 prefixes = __ballot_sync(FULL_MASK, flag == (u32)RR_FLAG_PREFIX);
                u32 take;  // lanes whose words are folded in
                if (prefixes) {
                    const u32 fp = __ffs(prefixes) - 1;
                    take = (fp == 31) ? FULL_MASK : ((2u << fp) - 1);
                    if (empties & take) continue;  // a closer tile has not published yet
                } else {
                    if (empties) continue;
                    take = FULL_MASK;
                }
                const bool mine = (take >> lane) & 1;
                const u32 h = mine ? (u32)((v >> 31) & 0x7fffffffu) : 0;
                const u32 kq = mine ? (u32)(v & 0x7fffffffu) : 0;
                exh = max(exh, warp_max(h));
                exk += warp_sum(kq);
                if (prefixes) break;
                t -= 32;
            }
            if (lane == 0) st_relaxed_u64(status + tile, rr_pack(RR_FLAG_PREFIX, max(exh, toth), exk + totk));
        }
        if (lane == 0) {
            s_exh = exh;
            s_exk = exk;
            if (totk) atomicAdd(&ctr->kept, totk);
        }
    }
    __syncthreads();
    const u32 exh = s_exh, exk = s_exk;

    // exclusive state in front of my first slot
    {
        // exclusive (over threads) versions of the warp scans
        u32 eh = __shfl_up_sync(FULL_MASK, ih, 1);
        u32 ek = __shfl_up_sync(FULL_MASK, ik, 1);
        if (lane == 0) { eh = 0; ek = 0; }
        u32 curh = max(exh, max(offh, eh));
        u32 curk = exk + offk + ek;
#pragma unroll
        for (int q = 0; q < RR_IPT; q++) {
            const u32 s = s0 + q;
            if (s < cnt) {
                const u32 j = base + s;
                if ((hbits >> q) & 1) curh = j + 1;
                const u32 jh = curh - 1;  // every slot has a head at or before it (slot gst[j] is one)
                const u32 g = __ldg(grp + j), gs = __ldg(gst + j);
                const u32 nr = g + (jh - gs);
                const u32 i = __ldg(idx + j);
                if (nr != g) rank[i] = nr;
                if ((kbits >> q) & 1) {
                    idx_out[curk] = i;
                    grp_out[curk] = nr;
                    gst_out[curk] = curk - (j - jh);
                    curk++;
                }
            }
        }
    }
}

// ---- emit -----------------------------------------------------------------------------------------
// out[rank[i]] = T[i-1] for every position that does not start a factor
__global__ void __launch_bounds__(256) k_emit(const u8 *__restrict__ T, u32 n, const u32 *__restrict__ rank,
                                              const u8 *__restrict__ flags, u8 *__restrict__ out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) return;
    out[ldg_stream_u32(rank + i)] = T[i - 1];
}
// factor heads receive the last byte of their own factor
__global__ void __launch_bounds__(256) k_emit_heads(const u8 *__restrict__ T, const u32 *__restrict__ FS, u32 F,
                                                    const u32 *__restrict__ rank, u8 *__restrict__ out)
{
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    out[rank[FS[f]]] = T[FS[f + 1] - 1];
}
// suffix-array variant: SA[rank[i]] = i
__global__ void __launch_bounds__(256) k_emit_sa(const u32 *__restrict__ rank, u32 n, i32 *__restrict__ sa)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sa[ldg_stream_u32(rank + i)] = (i32)i;
}

__global__ void k_set_u32(u32 *p, const u32 *idx_src, u32 value)
{
    p[*idx_src] = value;  // FS[F] = n
}
__global__ void __launch_bounds__(256) k_iota_u32(u32 *p, u32 n)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
