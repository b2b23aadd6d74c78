"""Executable specification (numpy / pure Python) of the PARALLEL algorithms the CUDA
kernels implement, kernel by kernel.  It exists so the maths can be checked against the
oracle on a CPU-only box; it is test code, not a product path (the product has no CPU
fallback and never imports this).

forward:  chunked Lyndon boundaries -> packed initial keys -> prefix doubling over the
          cyclic successor map with a shrinking live set -> emit
inverse:  stable LF map -> hashed splitters + sublist walks -> reduced-list min / sum
          jumping -> cycle offsets scan -> placement
"""
import numpy as np


# --------------------------------------------------------------------------- forward

def suffix_less(T, a, b):
    """T[a:] < T[b:] (a != b); the shorter suffix wins a tie."""
    n = len(T)
    while a < n and b < n:
        if T[a] != T[b]:
            return T[a] < T[b]
        a += 1
        b += 1
    return a == n  # a ran out first => a is the shorter => smaller


def lyndon_starts_chunked(T, B):
    """lyndon.cu: k_duval_chunks + k_chunk_min_scan + k_chunk_threshold."""
    n = len(T)
    flags = np.zeros(n, dtype=np.uint8)
    nch = (n + B - 1) // B
    last = np.zeros(nch, dtype=np.int64)
    # K1: every chunk runs Duval on the suffix T[b:], until the pending factor starts
    # at or beyond the chunk end; marks factor starts inside the chunk.
    for t in range(nch):
        b, e = t * B, min((t + 1) * B, n)
        f = b
        while f < e:
            i, k = f, f + 1
            settled = False
            while k < n:
                # past the chunk end only a repetition whose period still puts another copy
                # inside the chunk is undecided; otherwise nothing more can be marked
                if k >= e and f + (k - i) >= e:
                    settled = True
                    break
                if T[i] > T[k]:
                    break
                i = f if T[i] < T[k] else i + 1
                k += 1
            if settled:
                flags[f] = 1
                last[t] = f
                break
            p = k - i
            while f <= i and f < e:
                flags[f] = 1
                last[t] = f
                f += p
            if f <= i:  # remaining copies start beyond the chunk
                f = e
    # K2: exclusive prefix minimum (suffix order) of the per-chunk minima
    M = np.full(nch, -1, dtype=np.int64)
    run = -1
    for t in range(nch):
        M[t] = run
        if run < 0 or suffix_less(T, last[t], run):
            run = last[t]
    # K3: in each chunk the marked starts are decreasing in suffix order; keep the
    # suffix of that list that is below the minimum of everything before the chunk.
    for t in range(1, nch):
        b, e = t * B, min((t + 1) * B, n)
        cand = [p for p in range(b, e) if flags[p]]
        lo, hi = 0, len(cand)          # first index whose suffix < T[M[t]:]
        while lo < hi:
            mid = (lo + hi) // 2
            if suffix_less(T, cand[mid], M[t]):
                hi = mid
            else:
                lo = mid + 1
        for p in cand[:lo]:
            flags[p] = 0
    return np.flatnonzero(flags)


def forward(T, chunk=16, trace=None, tmax=0):
    """tmax > 0: groups of at most tmax members leave the rank-ordered live array for the
    text-ordered tuple set (forward.cuh: k_tuple_round / k_tuple_apply): members linked in a ring
    by text position, new rank = old rank + number of ring members with a smaller key2."""
    T = np.frombuffer(bytes(T), dtype=np.uint8)
    n = len(T)
    FS = np.append(lyndon_starts_chunked(T, chunk), n).astype(np.int64)
    fid = np.searchsorted(FS, np.arange(n), side="right") - 1
    fs, fl = FS[fid], FS[fid + 1] - FS[fid]
    lmax = int(fl.max())

    def succ_k(i, k):
        return fs[i] + (i - fs[i] + k) % fl[i]

    # alphabet compaction + packed initial key of k0 symbols (forward.cu: k_init_keys)
    present = np.zeros(256, dtype=bool)
    present[T] = True
    code = np.cumsum(present) - 1
    sigma = int(present.sum())
    bits = max(1, int(np.ceil(np.log2(sigma)))) if sigma > 1 else 1
    k0 = 64 // bits
    pos = np.arange(n)
    key = np.zeros(n, dtype=object)
    for _ in range(k0):
        key = key * (1 << bits) + code[T[pos]]
        pos = succ_k(pos, 1)
    order = np.array(sorted(range(n), key=lambda i: key[i]), dtype=np.int64)
    skey = key[order]

    rank = np.zeros(n, dtype=np.int64)
    # live arrays
    idx = order.copy()
    grp = np.zeros(n, dtype=np.int64)   # global rank of the group head, by live position
    gst = np.zeros(n, dtype=np.int64)   # live-array offset of the group start
    m = n
    k = k0
    rounds = 0
    NONE = -1
    ring = np.full(n, NONE, dtype=np.int64)   # tuple set: next member of my group, by text position
    tstats = dict(entered=0, t_rounds=0)

    def rerank(skey, idx, grp, gst, m, finalize=False):
        j = np.arange(m)
        head = (j == gst[:m])
        if finalize:
            head[:] = True
        else:
            head[1:] |= np.array([skey[a] != skey[a - 1] for a in range(1, m)], dtype=bool)
        nxt = np.append(head[1:], True)
        keep = ~(head & nxt)
        jh = np.maximum.accumulate(np.where(head, j, -1))
        newrank = grp[:m] + (jh - gst[:m])
        changed = newrank != grp[:m]
        rank[idx[:m][changed]] = newrank[changed]
        nheads = int(head.sum())
        # route: kept groups of at most tmax members go to the tuple set
        heads_pos = np.flatnonzero(head)
        sizes = np.diff(np.append(heads_pos, m))
        size_of = np.repeat(sizes, sizes)
        to_t = keep & (size_of <= tmax) if (tmax and not finalize) else np.zeros(m, dtype=bool)
        for h, sz in zip(heads_pos, sizes):
            if sz >= 2 and to_t[h]:
                members = idx[h:h + sz]
                ring[members] = np.roll(members, -1)
                tstats["entered"] += int(sz)
        stay = keep & ~to_t
        c = np.cumsum(stay) - stay       # exclusive
        nidx = idx[:m][stay]
        ngrp = newrank[stay]
        # offset of the group start in the compacted array: members of a group that stays are contiguous
        first_of_group = np.repeat(heads_pos, sizes)
        ngst = c[first_of_group][stay]
        kheads = int((head & stay).sum())
        return nidx, ngrp, ngst, int(stay.sum()), nheads, kheads

    def tuple_round(k, finalize=False):
        """phase A (k_tuple_round) computes on the old ranks / rings, phase B (k_tuple_apply) applies"""
        live = np.flatnonzero(ring != NONE)
        new_ring = ring.copy()
        dr = np.zeros(n, dtype=np.int64)
        split = False
        for i in live:
            ki = i if finalize else rank[succ_k(i, k)]
            less, eqn, mm = 0, NONE, ring[i]
            while mm != i:
                km = mm if finalize else rank[succ_k(mm, k)]
                less += km < ki
                split |= bool(km != ki)
                if km == ki and eqn == NONE:
                    eqn = mm
                mm = ring[mm]
            dr[i] = less
            new_ring[i] = NONE if finalize else eqn
        rank[live] += dr[live]
        ring[:] = new_ring
        return split, int((ring != NONE).sum())

    nidx, ngrp, ngst, m2, nheads, kheads = rerank(skey, idx, grp, gst, m)
    idx, grp, gst, m = nidx, ngrp, ngst, m2
    mt = int((ring != NONE).sum())
    while m + mt > 0 and k < 2 * lmax:
        groups_before = kheads
        split_t = False
        if mt:
            split_t, mt = tuple_round(k)
            tstats["t_rounds"] += 1
        nheads = groups_before
        if m:
            key2 = rank_before_t = None
        if m:
            # the rank-ordered sets gather AFTER the tuple set applied its new ranks (kernel order in
            # forward_core): groups are refined as a whole, so mixed depths are consistent
            key2 = rank[succ_k(idx[:m], k)]
            comp = gst[:m] * (n + 1) + key2
            perm = np.argsort(comp, kind="stable")
            skey = comp[perm]
            idx = idx[:m][perm]
            nidx, ngrp, ngst, m2, nheads, kheads2 = rerank(skey, idx, grp, gst, m)
        rounds += 1
        k *= 2
        if nheads == groups_before and not split_t:     # fixpoint: nothing split anywhere
            break
        if m:
            idx, grp, gst, m, kheads = nidx, ngrp, ngst, m2, kheads2
        mt = int((ring != NONE).sum())
    if m > 0:
        rerank(None, idx, grp, gst, m, finalize=True)
    if (ring != NONE).any():
        tuple_round(k, finalize=True)
    if trace is not None:
        trace.update(rounds=rounds, k=k, factors=len(FS) - 1, bits=bits, k0=k0, **tstats)
    # emit (forward.cu: k_emit / k_emit_heads)
    out = np.zeros(n, dtype=np.uint8)
    prevpos = np.arange(n) - 1
    isstart = np.zeros(n, dtype=bool)
    isstart[FS[:-1]] = True
    prevpos[isstart] = FS[fid[isstart] + 1] - 1
    out[rank] = T[prevpos]
    assert len(np.unique(rank)) == n
    return out.tobytes()


# --------------------------------------------------------------------------- inverse

def is_splitter(i, shift=26):
    return ((np.uint64(i) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)) >> np.uint64(shift) == 0


def inverse(Bs, shift=26):
    B = np.frombuffer(bytes(Bs), dtype=np.uint8)
    n = len(B)
    # LF map (inverse.cu: k_tile_hist, k_scan_*, k_lf_rank)
    cnt = np.bincount(B, minlength=256)
    C = np.cumsum(cnt) - cnt
    prev = np.zeros(n, dtype=np.int64)
    seen = np.zeros(256, dtype=np.int64)
    for i in range(n):
        prev[i] = C[B[i]] + seen[B[i]]
        seen[B[i]] += 1
    # splitters (k_splitter_*)
    spl = np.array([i for i in range(n) if is_splitter(i, shift)], dtype=np.int64)
    ns = len(spl)
    NONE = -1
    rec_head = np.full(n, NONE, dtype=np.int64)
    rec_off = np.zeros(n, dtype=np.int64)
    direct = np.zeros(n, dtype=bool)
    rec_head[spl] = np.arange(ns)
    nxt = np.zeros(ns, dtype=np.int64)
    w = np.zeros(ns, dtype=np.int64)
    mnv = np.zeros(ns, dtype=np.int64)
    mno = np.zeros(ns, dtype=np.int64)
    # walk (k_walk)
    for s in range(ns):
        i0 = spl[s]
        mn, mo, o = i0, 0, 1
        i = prev[i0]
        while not is_splitter(int(i), shift):
            rec_head[i], rec_off[i] = s, o
            if i < mn:
                mn, mo = i, o
            i = prev[i]
            o += 1
        nxt[s], w[s], mnv[s], mno[s] = rec_head[i], o, mn, mo
    # min jumping on the reduced list (k_min_jump)
    jmp, cm = nxt.copy(), mnv.copy()
    r = 0
    while (1 << r) < max(ns, 1):
        cm = np.minimum(cm, cm[jmp]) if ns else cm
        jmp = jmp[jmp] if ns else jmp
        r += 1
    if ns:
        cm = np.minimum(cm, cm[jmp])
    origin = mnv == cm
    # suffix sums with the origin as list terminal (k_sum_jump)
    ptr = np.where(origin[nxt], -1, nxt) if ns else nxt
    val = w.copy()
    for _ in range(r + 1):
        has = ptr >= 0
        val = np.where(has, val + val[np.where(has, ptr, 0)], val)
        ptr = np.where(has, ptr[np.where(has, ptr, 0)], -1)
    D = val
    lenAtMin = np.zeros(n, dtype=np.int64)
    cycL = np.zeros(n, dtype=np.int64)
    cycO = np.zeros(n, dtype=np.int64)
    for s in range(ns):
        if origin[s]:
            lenAtMin[cm[s]] = D[s]
            cycL[cm[s]] = D[s]
            cycO[cm[s]] = mno[s]
    # fallback: elements no walk reached (k_self_walk)
    dir_m = np.zeros(n, dtype=np.int64)
    dir_d = np.zeros(n, dtype=np.int64)
    for i in range(n):
        if rec_head[i] == NONE:
            j, steps, mn, mstep = prev[i], 1, i, 0
            while j != i:
                if j < mn:
                    mn, mstep = j, steps
                j = prev[j]
                steps += 1
            L = steps
            direct[i] = True
            dir_m[i], dir_d[i] = mn, (L - mstep) % L
            if mn == i:
                lenAtMin[i] = L
    off = np.cumsum(lenAtMin) - lenAtMin
    # per-splitter record (k_splitter_record)
    A = np.zeros(ns, dtype=np.int64)
    Ls = np.zeros(ns, dtype=np.int64)
    offs = np.zeros(ns, dtype=np.int64)
    for s in range(ns):
        L = cycL[cm[s]]
        P = L - D[s]
        A[s] = (P - cycO[cm[s]]) % L
        Ls[s] = L
        offs[s] = off[cm[s]]
    # placement (k_place)
    out = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        if direct[i]:
            pos = n - 1 - off[dir_m[i]] - dir_d[i]
        else:
            s = rec_head[i]
            d = A[s] + rec_off[i]
            if d >= Ls[s]:
                d -= Ls[s]
            pos = n - 1 - offs[s] - d
        out[pos] = B[i]
    return out.tobytes()


def inverse_staged(Bs, shift=26, slot=256, Q=64):
    """inverse.cuh, staged path: k_inv_spl_write (blkoff) -> k_inv_walk_stage (warps of 32 lanes with
    lane refill over a range of Q sublists, bytes parked in 32-byte sectors of a slot, cont[] at
    offset `slot`) -> k_inv_resolve_next (sid_of) -> list ranking as before -> k_inv_place_copy +
    k_inv_walk_tail.  Mirrors the CUDA control flow statement by statement."""
    B = np.frombuffer(bytes(Bs), dtype=np.uint8)
    n = len(B)
    cnt = np.bincount(B, minlength=256)
    C = np.concatenate((np.cumsum(cnt) - cnt, [n]))  # C[256] = n

    def byte_of_rank(p):
        c, step = 0, 128
        while step:
            if C[c + step] <= p:
                c += step
            step >>= 1
        return c

    prev = np.zeros(n, dtype=np.int64)
    seen = np.zeros(256, dtype=np.int64)
    for i in range(n):
        prev[i] = C[B[i]] + seen[B[i]]
        seen[B[i]] += 1
    spl = [i for i in range(n) if is_splitter(i, shift)]
    ns = len(spl)
    blkoff = np.zeros((n >> 6) + 2, dtype=np.int64)
    run = 0
    for i in range(n):
        if i % 64 == 0:
            blkoff[i >> 6] = run
        run += bool(is_splitter(i, shift))

    def sid_of(p):
        s = blkoff[p >> 6]
        for q in range(p & ~63, p):
            s += bool(is_splitter(q, shift))
        return int(s)

    stage = np.full(ns * slot, 0xEE, dtype=np.uint8)  # poison: unwritten bytes must never be placed
    nxt = np.zeros(ns, dtype=np.int64)
    wlen = np.zeros(ns, dtype=np.int64)
    mnv = np.zeros(ns, dtype=np.int64)
    mno = np.zeros(ns, dtype=np.int64)
    cont = np.full(ns, -1, dtype=np.int64)
    visited = np.zeros(n, dtype=bool)
    total = 0
    NONE = -1
    nwarps = (ns + Q - 1) // Q
    for w in range(nwarps):
        nxt_s, hi = w * Q, min(ns, (w + 1) * Q)
        st = [dict(s=NONE) for _ in range(32)]
        while True:
            idle = [l for l in range(32) if st[l]["s"] == NONE]
            if idle:
                for r, l in enumerate(idle):
                    cand = nxt_s + r
                    if cand < hi:
                        st[l] = dict(s=cand, i=spl[cand], o=0, mn=spl[cand], mo=0, acc=bytearray(32))
                nxt_s += len(idle)
                if all(x["s"] == NONE for x in st):
                    break
            for x in st:
                if x["s"] == NONE:
                    continue
                s_, i, o = x["s"], x["i"], x["o"]
                p = int(prev[i])
                visited[i] = True
                if o < slot:
                    b = o & 31
                    x["acc"][b] = byte_of_rank(p)
                    if b == 31:
                        base = s_ * slot + (o & ~31)
                        stage[base:base + 32] = np.frombuffer(bytes(x["acc"]), np.uint8)
                        x["acc"] = bytearray(32)
                elif o == slot:
                    cont[s_] = i
                o += 1
                x["o"] = o
                if is_splitter(p, shift):
                    if o <= slot and (o & 31):
                        base = s_ * slot + ((o - 1) & ~31)
                        stage[base:base + 32] = np.frombuffer(bytes(x["acc"]), np.uint8)
                    nxt[s_], wlen[s_], mnv[s_], mno[s_] = p, o, x["mn"], x["mo"]
                    total += o
                    x["s"] = NONE
                else:
                    x["i"] = p
                    if p < x["mn"]:
                        x["mn"], x["mo"] = p, o
    nxt = np.array([sid_of(int(p)) for p in nxt], dtype=np.int64)
    w_ = wlen
    # reduced list: identical to inverse()
    jmp, cm = nxt.copy(), mnv.copy()
    r = 0
    while (1 << r) < max(ns, 1):
        cm = np.minimum(cm, cm[jmp]) if ns else cm
        jmp = jmp[jmp] if ns else jmp
        r += 1
    if ns:
        cm = np.minimum(cm, cm[jmp])
    origin = mnv == cm
    ptr = np.where(origin[nxt], -1, nxt) if ns else nxt
    val = w_.copy()
    for _ in range(r + 1):
        has = ptr >= 0
        val = np.where(has, val + val[np.where(has, ptr, 0)], val)
        ptr = np.where(has, ptr[np.where(has, ptr, 0)], -1)
    D = val
    lenAtMin = np.zeros(n, dtype=np.int64)
    cycL = np.zeros(n, dtype=np.int64)
    cycO = np.zeros(n, dtype=np.int64)
    for s_ in range(ns):
        if origin[s_]:
            lenAtMin[cm[s_]] = D[s_]
            cycL[cm[s_]] = D[s_]
            cycO[cm[s_]] = mno[s_]
    direct = np.zeros(n, dtype=bool)
    dir_m = np.zeros(n, dtype=np.int64)
    dir_d = np.zeros(n, dtype=np.int64)
    if total != n:
        for i in range(n):
            if not visited[i]:
                j, steps, mn, mstep = prev[i], 1, i, 0
                while j != i:
                    if j < mn:
                        mn, mstep = j, steps
                    j = prev[j]
                    steps += 1
                direct[i] = True
                dir_m[i], dir_d[i] = mn, (steps - mstep) % steps
                if mn == i:
                    lenAtMin[i] = steps
    off = np.cumsum(lenAtMin) - lenAtMin
    out = np.full(n, 0xDD, dtype=np.uint8)
    written = np.zeros(n, dtype=np.int64)
    for s_ in range(ns):
        L = int(cycL[cm[s_]])
        A = int((L - D[s_] - cycO[cm[s_]]) % L)
        top = n - 1 - int(off[cm[s_]])
        # k_inv_place_copy
        for o in range(min(int(wlen[s_]), slot)):
            d = A + o
            if d >= L:
                d -= L
            out[top - d] = stage[s_ * slot + o]
            written[top - d] += 1
        # k_inv_walk_tail
        if wlen[s_] > slot:
            d = A + slot
            if d >= L:
                d -= L
            i = int(cont[s_])
            while True:
                p = int(prev[i])
                out[top - d] = byte_of_rank(p)
                written[top - d] += 1
                d += 1
                if d == L:
                    d = 0
                i = p
                if is_splitter(i, shift):
                    break
    for i in range(n):
        if direct[i]:
            pos = n - 1 - off[dir_m[i]] - dir_d[i]
            out[pos] = B[i]
            written[pos] += 1
    assert (written == 1).all(), "every output byte is written exactly once"
    return out.tobytes()


# ---- models of the round-1b kernels --------------------------------------------------------------

def owned_ranges(gst, T):
    """Ownership rule of k_local_sort_warp (T = 32) and k_local_sort_cta (T = 4096): worker c owns
    the whole groups between the group holding live slot c*T and the group holding slot (c+1)*T.
    gst[j] = live-array offset at which the group of slot j starts.  Returns [(lo, hi)]."""
    m = len(gst)
    out = []
    for c in range((m + T - 1) // T):
        lo = int(gst[c * T])
        hi = int(gst[(c + 1) * T]) if (c + 1) * T < m else m
        out.append((lo, hi))
    return out


def cta_sort_key(g_local, r, slot):
    """64-bit word ordered by the bitonic network of k_local_sort_cta: local group start (13 bits) |
    key2 (31 bits) | slot (13 bits)."""
    return (int(g_local) << 44) | (int(r) << 13) | int(slot)


def bitonic_sort(a):
    """the compare-exchange schedule of k_local_sort_cta (strides >= 8 in shared memory, smaller ones
    in registers -- the order of the exchanges is the same) on a power-of-two list"""
    a = list(a)
    P = len(a)
    k2 = 2
    while k2 <= P:
        j = k2 >> 1
        while j > 0:
            for t in range(P // 2):
                i = ((t & ~(j - 1)) << 1) | (t & (j - 1))
                l = i + j
                up = (i & k2) == 0
                if (a[i] > a[l]) == up:
                    a[i], a[l] = a[l], a[i]
            j >>= 1
        k2 <<= 1
    return a


def binned_scatter(pos, val, n, kb):
    """first re-rank / large emit: pairs binned (stably) by the top 8 bits of the target position,
    then scattered bin by bin; must equal the direct scatter"""
    shift = kb - 8
    order = np.argsort(pos >> shift, kind="stable")
    out = np.zeros(n, dtype=val.dtype)
    out[pos[order]] = val[order]
    return out


def binned_scatter_counted(pos, val, rank, kb):
    """later re-ranks: only the live positions are written, the bin starts come from a count of the positions per
    region (k_bin_count + scan) instead of the closed form; untouched entries of rank keep their value"""
    shift = kb - 8
    counts = np.bincount(pos >> shift, minlength=256)
    base = np.concatenate(([0], np.cumsum(counts)[:-1]))
    bin_pos = np.empty_like(pos)
    bin_val = np.empty_like(val)
    fill = base.copy()
    for p, v in zip(pos, val):          # the onesweep pass: stable inside a bin
        b = p >> shift
        bin_pos[fill[b]] = p
        bin_val[fill[b]] = v
        fill[b] += 1
    out = rank.copy()
    out[bin_pos] = bin_val              # k_scatter_pairs, region by region
    return out


def onesweep_tile_permutation(digits):
    """k_onesweep_pass: sorted slot -> raw position of one tile (stable by digit), built the way the
    kernel does: per-warp ranks in slot order, exclusive warp offsets per digit, digit starts"""
    TILE = len(digits)
    inv = np.full(TILE, -1, dtype=np.int64)
    counts = np.bincount(digits, minlength=256)
    dstart = np.concatenate(([0], np.cumsum(counts)[:-1]))
    seen = np.zeros(256, dtype=np.int64)
    # warp-striped slots visit raw positions in increasing order per warp and the warps' counters are
    # offset by the earlier warps' totals, so the net effect is "stable by digit in raw order"
    for p in range(TILE):
        d = digits[p]
        inv[dstart[d] + seen[d]] = p
        seen[d] += 1
    return inv


def initial_keys(codes, starts, bits, k0, extra=0):
    """k_init_keys: key[i] = the first k0 symbols (codes, `bits` each) of the rotation of i's factor starting at i,
    followed by the top `extra` bits of symbol k0 + 1 (extra < bits); `starts` = factor starts + [n]"""
    n = len(codes)
    keys = [0] * n
    for f in range(len(starts) - 1):
        s, e = starts[f], starts[f + 1]
        L = e - s
        for i in range(s, e):
            k = 0
            for t in range(k0):
                k = (k << bits) | int(codes[s + (i - s + t) % L])
            if extra:
                k = (k << extra) | (int(codes[s + (i - s + k0) % L]) >> (bits - extra))
            keys[i] = k
    return keys


def digit_hists_from_windows(keys, bits, k0, extra=0):
    """k_init_keys' window histogram + k_digit_hists: the 256-bin histogram of every radix digit of the keys, read
    off ONE histogram of the leading `wsyms` symbols.  Returns None where the driver keeps k_radix_hist
    (windows wider than 14 bits)."""
    k0p, d = k0 + (1 if extra else 0), (bits - extra) if extra else 0
    P0 = -(-(k0 * bits + extra) // 8)
    span = []
    for p in range(P0):
        lo_bit, hi_bit = 8 * p + d, min(8 * p + 7 + d, k0p * bits - 1)
        span.append(hi_bit // bits - lo_bit // bits + 1)
    wsyms = max(span)
    if wsyms * bits > 14 or k0 < wsyms:
        return None
    wshift = extra + bits * (k0 - wsyms)
    whist = np.bincount(np.array([k >> wshift for k in keys], dtype=np.int64), minlength=1 << (wsyms * bits))
    out = np.zeros((P0, 256), dtype=np.int64)
    for p in range(P0):
        lo_bit, hi_bit = 8 * p + d, min(8 * p + 7 + d, k0p * bits - 1)
        t_lo, t_hi = k0p - 1 - lo_bit // bits, k0p - 1 - hi_bit // bits
        wp = t_lo - t_hi + 1
        drop = bits * (wsyms - wp)
        sh_r = lo_bit - bits * (k0p - 1 - t_lo)
        for w in np.flatnonzero(whist):
            out[p][((int(w) >> drop) >> sh_r) & 255] += whist[w]
    return out
