// lyndon.cuh -- parallel Lyndon-boundary kernels.
//
// The reference finds factor starts as the strict prefix minima of the inverse suffix
// array (/root/reference/mk_bwts_sa.c:126-129), i.e. position i starts a factor iff the
// suffix T[i..n) is smaller than every earlier suffix.  Here, with no suffix array:
//   1. k_duval_chunks   one thread per chunk [b,e) runs Duval's algorithm on the suffix
//                       T[b..n) until the pending factor starts at or after e.  That marks
//                       exactly the positions of the chunk that are smaller than every
//                       earlier suffix *starting in the same chunk* (a decreasing list).
//   2. k_chunkmin_*     prefix minimum, in suffix order, of the chunks' last marks (= each
//                       chunk's smallest suffix): per-group reduce, Hillis-Steele levels.
//   3. k_chunk_threshold a mark survives iff its suffix is below that minimum; survivors
//                       are a tail of the chunk's list (binary search, warp-wide compares).
#pragma once
#include "common.cuh"

#define LY_GROUP 32  // chunks per warp in the min-scan kernels

// Work budget of the chunk kernels.  Periodic texts (a^n, a long run of zeros, ...) make
// the comparisons below run to the end of the periodic stretch for every chunk, which is
// quadratic; when a thread / warp has spent more than its budget it raises *abort and every
// kernel winds down quickly.  The host then takes the factor starts from a suffix sort
// instead (strict prefix minima of the inverse suffix array, the reference's own criterion).
struct LyBudget {
    u32 *abort;      // global flag
    u32 limit;       // per warp: KiB of comparison; Duval: 8-byte steps past the chunk, all threads together
    u32 *counter;    // Duval only: global count of such steps
};

// T[a..n) < T[b..n) for a != b; whole warp cooperates (uniform arguments and result).
// `spent` accumulates KiB compared by this warp; gives up (result meaningless) once over budget.
static __device__ __forceinline__ bool suffix_less_warp(const u8 *__restrict__ T, u32 n, u32 a, u32 b,
                                                        const LyBudget &bud, u32 &spent)
{
    const u32 lane = lane_id();
    u32 off = 0;
    for (;;) {
        // every lane checks 32 bytes (4 independent 8-byte compares), the warp 1 KiB per step
        const u32 pa = a + off + lane * 32, pb = b + off + lane * 32;
        u32 first = 4;  // index of my first differing (or unsafe) 8-byte word
        if ((u64)pa + 40 <= n && (u64)pb + 40 <= n) {
            u64 x[4], y[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { x[q] = load8_unaligned(T + pa + 8 * q); y[q] = load8_unaligned(T + pb + 8 * q); }
#pragma unroll
            for (int q = 3; q >= 0; q--) if (x[q] != y[q]) first = q;
        } else {
            first = 0;  // too close to the end: let the byte loop decide
        }
        const u32 mask = __ballot_sync(FULL_MASK, first < 4);
        if (mask == 0) {
            off += 1024;
            spent++;
            if ((spent & 63) == 0) {
                u32 stop = 0;
                if (lane == 0) {
                    if (spent > bud.limit) atomicExch(bud.abort, 1u);
                    stop = *(volatile u32 *)bud.abort;
                }
                if (__shfl_sync(FULL_MASK, stop, 0)) return false;
            }
            continue;
        }
        const u32 l = __ffs(mask) - 1;
        const u32 w = __shfl_sync(FULL_MASK, first, l);
        u32 p = a + off + l * 32 + w * 8, q = b + off + l * 32 + w * 8;
        while (p < n && q < n) {
            const u8 ca = T[p], cb = T[q];
            if (ca != cb) return ca < cb;
            p++; q++;
        }
        return p >= n;  // the suffix that ends first is the smaller one
    }
}

// ---- 1. Duval per chunk -----------------------------------------------------------------
// One thread per chunk, written as ONE flat loop over a small state machine (scan a byte / skip 8 bytes of
// a repetition / emit a factor start): the nested loops of the textbook form put the 32 lanes of a warp,
// which run 32 different scans, at 32 different places of the code -- ncu on the nested form: 92 % issue
// activity for 3 % DRAM, 6.6 G warp instructions for 1 Gi bytes, 16x what converged lanes need.
__global__ void __launch_bounds__(128) k_duval_chunks(const u8 *__restrict__ T, u32 n, u32 chunk, u32 nch,
                                                      u8 *__restrict__ flags, u32 *__restrict__ chunk_last,
                                                      LyBudget bud)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nch) return;
    u32 spent = 0;  // 8-byte steps taken past the end of the chunk
    const u32 b = t * chunk;
    const u32 e = min(n, b + chunk);
    u32 f = b, last = b, i = b, k = b + 1, p = 0;
    enum { SCAN = 0, EMIT = 1, DONE = 2 };
    int mode = (f < e) ? SCAN : DONE;
    bool rep = false;  // 16 bytes into a repetition: compare 8 bytes at a time
    while (mode != DONE) {
        if (mode == SCAN) {
            if (rep) {
                if ((u64)k + 16 <= n && load8_unaligned(T + i) == load8_unaligned(T + k)) {
                    i += 8; k += 8;
                    if (k >= e && ((++spent) & 1023) == 0) {
                        if (atomicAdd(bud.counter, 1024u) > bud.limit) atomicExch(bud.abort, 1u);
                        if (*(volatile u32 *)bud.abort) mode = DONE;
                    }
                } else {
                    rep = false;
                }
            } else if (k >= n) {
                p = k - i;
                mode = EMIT;
            } else if (k >= e && f + (k - i) >= e) {
                // past the chunk end only a repetition with a period short enough to put another copy inside
                // the chunk is still undecided; anything else cannot mark more starts: T[f..] is one pending
                // word reaching beyond the chunk, f is its only start
                flags[f] = 1;
                last = f;
                mode = DONE;
            } else {
                const u8 ci = T[i], ck = T[k];
                if (ci > ck) {
                    p = k - i;
                    mode = EMIT;
                } else {
                    const bool eq = ci == ck;
                    i = eq ? i + 1 : f;
                    k++;
                    rep = eq && (i - f >= 16);
                }
            }
        } else {  // EMIT: every copy of the word of period p that starts at or before i starts a factor
            if (f <= i && f < e) {
                flags[f] = 1;
                last = f;
                f += p;
            } else {
                if (f <= i) f = e;  // the remaining copies start beyond this chunk
                if (f < e) { i = f; k = f + 1; rep = false; mode = SCAN; } else mode = DONE;
            }
        }
    }
    chunk_last[t] = last;
}

// ---- 2. exclusive prefix minimum of chunk_last in suffix order ---------------------------
static __device__ __forceinline__ u32 suffix_min_warp(const u8 *T, u32 n, u32 a, u32 b, const LyBudget &bud,
                                                      u32 &spent)
{
    if (a == NONE32) return b;
    if (b == NONE32) return a;
    return suffix_less_warp(T, n, b, a, bud, spent) ? b : a;
}

// one warp per group of LY_GROUP chunks
__global__ void __launch_bounds__(128) k_chunkmin_reduce(const u8 *__restrict__ T, u32 n,
                                                         const u32 *__restrict__ chunk_last, u32 nch,
                                                         u32 *__restrict__ group_min, u32 ngroups, LyBudget bud)
{
    const u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= ngroups) return;
    const u32 lo = g * LY_GROUP, hi = min(nch, lo + LY_GROUP);
    u32 run = NONE32, spent = 0;
    for (u32 t = lo; t < hi; t++) run = suffix_min_warp(T, n, run, chunk_last[t], bud, spent);
    if (lane_id() == 0) group_min[g] = run;
}

// one Hillis-Steele level of the inclusive prefix minimum over the groups (warp per group):
// out[g] = min(in[g], in[g - stride]).  ceil(log2(ngroups)) launches, ping-pong buffers.
__global__ void __launch_bounds__(128) k_chunkmin_level(const u8 *__restrict__ T, u32 n, const u32 *__restrict__ in,
                                                        u32 *__restrict__ out, u32 ngroups, u32 stride, LyBudget bud)
{
    const u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= ngroups) return;
    if (*(volatile u32 *)bud.abort) return;  // an earlier level ran out of budget: the scan is abandoned
    u32 v = in[g], spent = 0;
    if (g >= stride) v = suffix_min_warp(T, n, in[g - stride], v, bud, spent);
    if (lane_id() == 0) out[g] = v;
}

// ---- 2b. the same scan, work-efficient, with CTA-wide comparisons above the first level -------
// The Hillis-Steele levels above compare ngroups * log2(ngroups) pairs of suffixes; on periodic
// inputs (the 64 KiB-tiled C3 text: chunk minima of different tile copies agree for up to 1 MiB)
// every comparison streams megabytes and the levels cost 30 ms of a 36 ms boundary search.  Here:
// reduce 32 -> 1 level by level (k_sufmin_reduce_cta), then hand the exclusive prefixes back down
// (k_sufmin_down_cta): ~2 * ngroups comparisons in all, each done by a whole CTA (32 KiB per step),
// because the upper levels have few groups and a single warp would crawl through a 1 MiB match.
#define LY_CTA 1024
struct LyCtaShared {
    u32 cand[2][LY_CTA / 32];
    u32 res, stop;
};
// T[a..n) < T[b..n), a != b; all LY_CTA threads of the block call it with the same arguments.
static __device__ bool suffix_less_cta(const u8 *__restrict__ T, u32 n, u32 a, u32 b, LyCtaShared &sh,
                                       const LyBudget &bud, u32 &spent)
{
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u64 off = 0;
    for (u32 it = 0;; it++) {
        const u64 pa = (u64)a + off + (u64)tid * 32, pb = (u64)b + off + (u64)tid * 32;
        u32 first = 4;  // index of my first differing (or unsafe) 8-byte word
        if (pa + 40 <= n && pb + 40 <= n) {
            u64 x[4], y[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { x[q] = load8_unaligned(T + pa + 8 * q); y[q] = load8_unaligned(T + pb + 8 * q); }
#pragma unroll
            for (int q = 3; q >= 0; q--) if (x[q] != y[q]) first = q;
        } else {
            first = 0;  // too close to the end (or past it): the byte loop decides
        }
        const u32 mine = warp_min(first < 4 ? tid : NONE32);
        u32 *cand = sh.cand[it & 1];
        if (lane == 0) cand[warp] = mine;
        if (tid == 0 && (it & 63) == 63) {
            spent += 64 * (LY_CTA * 32 / 1024);  // KiB compared since the last check
            if (spent > bud.limit) atomicExch(bud.abort, 1u);
            sh.stop = *(volatile u32 *)bud.abort;
        }
        __syncthreads();
        u32 best = (lane < LY_CTA / 32) ? cand[lane] : NONE32;
        best = warp_min(best);
        if (sh.stop) return false;  // over budget: the result is meaningless, the caller winds down
        if (best == NONE32) { off += (u64)LY_CTA * 32; continue; }
        if (tid == best) {
            u64 p = pa + 8 * first, q = pb + 8 * first;
            if (p > n) p = n;  // a window that starts past the end
            if (q > n) q = n;
            bool less;
            for (;;) {
                if (p >= n || q >= n) { less = p >= n; break; }  // the suffix that ends first is the smaller one
                const u8 ca = T[p], cb = T[q];
                if (ca != cb) { less = ca < cb; break; }
                p++; q++;
            }
            sh.res = less ? 1u : 0u;
        }
        __syncthreads();
        return sh.res != 0;
    }
}
static __device__ __forceinline__ u32 suffix_min_cta(const u8 *T, u32 n, u32 a, u32 b, LyCtaShared &sh, const LyBudget &bud,
                                                     u32 &spent)
{
    if (a == NONE32) return b;
    if (b == NONE32) return a;
    const bool less = suffix_less_cta(T, n, b, a, sh, bud, spent);
    __syncthreads();  // sh.res / sh.cand are reused by the next comparison
    return less ? b : a;
}

// out[g] = smallest suffix among in[32 g .. 32 g + 31]   (one CTA per g)
__global__ void __launch_bounds__(LY_CTA) k_sufmin_reduce_cta(const u8 *__restrict__ T, u32 n, const u32 *__restrict__ in,
                                                              u32 nin, u32 *__restrict__ out, LyBudget bud)
{
    __shared__ LyCtaShared sh;
    if (threadIdx.x == 0) sh.stop = 0;
    __syncthreads();
    const u32 g = blockIdx.x, lo = g * LY_GROUP, hi = min(nin, lo + LY_GROUP);
    u32 run = NONE32, spent = 0;
    for (u32 t = lo; t < hi; t++) run = suffix_min_cta(T, n, run, in[t], sh, bud, spent);
    if (threadIdx.x == 0) out[g] = run;
}
// excl[32 g + j] = smallest suffix among (everything before group g) and in[32 g .. 32 g + j - 1];
// parent_excl == nullptr: nothing lies before group 0 (top level, one CTA)
__global__ void __launch_bounds__(LY_CTA) k_sufmin_down_cta(const u8 *__restrict__ T, u32 n, const u32 *__restrict__ in,
                                                            u32 nin, const u32 *__restrict__ parent_excl,
                                                            u32 *__restrict__ excl, LyBudget bud)
{
    __shared__ LyCtaShared sh;
    if (threadIdx.x == 0) sh.stop = 0;
    __syncthreads();
    const u32 g = blockIdx.x, lo = g * LY_GROUP, hi = min(nin, lo + LY_GROUP);
    u32 run = parent_excl ? parent_excl[g] : NONE32, spent = 0;
    for (u32 t = lo; t < hi; t++) {
        if (threadIdx.x == 0) excl[t] = run;
        if (t + 1 < hi) run = suffix_min_cta(T, n, run, in[t], sh, bud, spent);
    }
}

// ---- 3. per chunk: drop the marks that are not below the minimum of everything before ----
// position of the j-th (0-based) mark in [b,e); every lane owns a contiguous slice.
static __device__ u32 select_mark_warp(const u8 *flags, u32 b, u32 e, u32 j)
{
    const u32 lane = lane_id();
    const u32 len = e - b, per = (len + 31) / 32;
    const u32 lo = min(e, b + lane * per), hi = min(e, lo + per);
    u32 c = 0;
    for (u32 p = lo; p < hi; p++) c += flags[p];
    const u32 incl = warp_incl_sum(c), excl = incl - c;
    const u32 owner = __ffs(__ballot_sync(FULL_MASK, j < incl)) - 1;
    u32 pos = NONE32;
    if (lane == owner) {
        u32 need = j - excl;
        for (u32 p = lo; p < hi; p++)
            if (flags[p]) { if (need == 0) { pos = p; break; } need--; }
    }
    return __shfl_sync(FULL_MASK, pos, owner);
}

static __device__ void clear_flags_warp(u8 *flags, u32 lo, u32 hi)
{
    for (u32 p = lo + lane_id(); p < hi; p += 32) flags[p] = 0;
}

__global__ void __launch_bounds__(128) k_chunk_threshold(const u8 *__restrict__ T, u32 n, u32 chunk, u32 nch,
                                                         u8 *__restrict__ flags, const u32 *__restrict__ chunk_last,
                                                         const u32 *__restrict__ group_pre, int pre_is_exclusive,
                                                         u32 ngroups, LyBudget bud)
{
    const u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= ngroups) return;
    const u32 lo = g * LY_GROUP, hi = min(nch, lo + LY_GROUP);
    // smallest suffix of everything before this group: exclusive prefixes (hierarchical scan) or the
    // inclusive ones of the Hillis-Steele levels
    u32 run = pre_is_exclusive ? group_pre[g] : (g ? group_pre[g - 1] : NONE32);
    u32 spent = 0;
    for (u32 t = lo; t < hi; t++) {
        const u32 mlast = chunk_last[t];
        if (run != NONE32) {
            const u32 b = t * chunk, e = min(n, b + chunk);
            {
                u32 stop = 0;
                if (lane_id() == 0) stop = *(volatile u32 *)bud.abort;
                if (__shfl_sync(FULL_MASK, stop, 0)) return;  // warp-uniform exit
            }
            if (!suffix_less_warp(T, n, mlast, run, bud, spent)) {
                clear_flags_warp(flags, b, e);  // nothing in this chunk beats the earlier minimum
            } else if (!suffix_less_warp(T, n, b, run, bud, spent)) {
                // marks are decreasing in suffix order: first (=b) fails, last succeeds
                u32 c = 0;
                {
                    const u32 per = (e - b + 31) / 32;
                    const u32 l0 = min(e, b + lane_id() * per), h0 = min(e, l0 + per);
                    for (u32 p = l0; p < h0; p++) c += flags[p];
                    c = warp_sum(c);
                }
                u32 a = 0, z = c - 1;  // invariant: mark a fails, mark z succeeds
                u32 zpos = mlast;
                while (z - a > 1) {
                    const u32 mid = (a + z) >> 1;
                    const u32 pos = select_mark_warp(flags, b, e, mid);
                    if (suffix_less_warp(T, n, pos, run, bud, spent)) { z = mid; zpos = pos; } else a = mid;
                }
                clear_flags_warp(flags, b, zpos);
            }
            __syncwarp();
        }
        run = suffix_min_warp(T, n, run, mlast, bud, spent);
    }
}

// ---- fallback: factor starts = strict prefix minima of the inverse suffix array ---------------
// (the reference's criterion, /root/reference/mk_bwts_sa.c:126-129).  isa = ranks of a suffix sort.
#define PM_TILE 4096
__global__ void __launch_bounds__(256) k_tile_min_u32(const u32 *__restrict__ isa, u32 n, u32 *__restrict__ tile_min)
{
    __shared__ u32 ws[8];
    const u32 base = blockIdx.x * PM_TILE;
    u32 v = NONE32;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const u32 i = base + q * 256 + threadIdx.x;
        if (i < n) v = min(v, isa[i]);
    }
    v = warp_min(v);
    if (lane_id() == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = NONE32;
        for (int w = 0; w < 8; w++) t = min(t, ws[w]);
        tile_min[blockIdx.x] = t;
    }
}
// single block: exclusive prefix minimum of the tile minima, in place
__global__ void __launch_bounds__(1024) k_tile_min_scan(u32 *__restrict__ tile_min, u32 ntiles)
{
    __shared__ u32 part[1024];
    const u32 per = (ntiles + 1023) / 1024;
    const u32 lo = min(ntiles, threadIdx.x * per), hi = min(ntiles, lo + per);
    u32 v = NONE32;
    for (u32 t = lo; t < hi; t++) v = min(v, tile_min[t]);
    part[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = NONE32;
        for (u32 t = 0; t < 1024; t++) { const u32 x = part[t]; part[t] = run; run = min(run, x); }
    }
    __syncthreads();
    u32 run = part[threadIdx.x];
    for (u32 t = lo; t < hi; t++) { const u32 x = tile_min[t]; tile_min[t] = run; run = min(run, x); }
}
// flags[i] = isa[i] < min(isa[0..i))   (thread owns 16 consecutive positions)
__global__ void __launch_bounds__(256) k_prefix_min_flags(const u32 *__restrict__ isa, u32 n,
                                                          const u32 *__restrict__ tile_excl, u8 *__restrict__ flags)
{
    __shared__ u32 ws[8];
    const u32 base = blockIdx.x * PM_TILE + threadIdx.x * 16;
    u32 v[16];
    u32 mine = NONE32;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        v[q] = (base + q < n) ? isa[base + q] : NONE32;
        mine = min(mine, v[q]);
    }
    // exclusive prefix minimum over the threads of the block
    u32 incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane_id() >= (u32)o) incl = min(incl, y);
    }
    if (lane_id() == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 run = __shfl_up_sync(FULL_MASK, incl, 1);
    if (lane_id() == 0) run = NONE32;
    run = min(run, tile_excl[blockIdx.x]);
    for (u32 w = 0; w < (threadIdx.x >> 5); w++) run = min(run, ws[w]);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        if (base + q < n) flags[base + q] = (v[q] < run) ? 1 : 0;  // position 0: run == NONE32
        run = min(run, v[q]);
    }
}

// ---- factor table ------------------------------------------------------------------------
// tile = 4096 flags per block of 256 threads (16 bytes per thread)
#define FL_TILE 4096
__global__ void __launch_bounds__(256) k_flag_count(const u8 *__restrict__ flags, u32 n, u32 *__restrict__ tile_cnt)
{
    __shared__ u32 ws[8];
    const u32 base = blockIdx.x * FL_TILE + threadIdx.x * 16;
    u32 c = 0;
    if (base + 16 <= n) {
        const uint4 v = *(const uint4 *)(flags + base);
        c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);  // flags are 0/1 bytes
    } else {
        for (u32 p = base; p < n; p++) c += flags[p];
    }
    c = warp_sum(c);
    if (lane_id() == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = 0;
        for (int w = 0; w < 8; w++) s += ws[w];
        tile_cnt[blockIdx.x] = s;
    }
}

// tile_off = exclusive scan of tile_cnt.  Writes FS[rank] = position for every mark.
__global__ void __launch_bounds__(256) k_flag_write(const u8 *__restrict__ flags, u32 n,
                                                    const u32 *__restrict__ tile_off, u32 *__restrict__ FS)
{
    __shared__ u32 ws[8];
    const u32 base = blockIdx.x * FL_TILE + threadIdx.x * 16;
    u8 loc[16];
    u32 c = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        loc[q] = (base + q < n) ? flags[base + q] : 0;
        c += loc[q];
    }
    const u32 incl = warp_incl_sum(c);
    if (lane_id() == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 off = tile_off[blockIdx.x] + incl - c;
    for (u32 w = 0; w < (threadIdx.x >> 5); w++) off += ws[w];
#pragma unroll
    for (int q = 0; q < 16; q++)
        if (loc[q]) FS[off++] = base + q;
}

// longest factor + coarse index.  FS[F] = n must already be in place.
__global__ void k_factor_lmax(const u32 *__restrict__ FS, u32 F, u32 *__restrict__ lmax)
{
    u32 best = 0;
    for (u32 f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x)
        best = max(best, FS[f + 1] - FS[f]);
    best = warp_max(best);
    if (lane_id() == 0 && best) atomicMax(lmax, best);
}

__global__ void k_coarse_index(const u32 *__restrict__ FS, u32 F, u32 *__restrict__ cidx, u32 nblk)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblk) return;
    if (b == nblk) { cidx[b] = F - 1; return; }
    const u32 pos = b << COARSE_BITS;
    u32 lo = 0, hi = F - 1;  // largest f with FS[f] <= pos
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if (FS[mid] <= pos) lo = mid; else hi = mid - 1;
    }
    cidx[b] = lo;
}
