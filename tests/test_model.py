"""CPU test: the executable specification of the parallel algorithms (the same maths the
CUDA kernels implement, kernel by kernel) agrees with the oracle."""
import numpy as np

import helpers
import model_gpu_algorithms as model


def _cases():
    rng = np.random.default_rng(1)
    out = []
    for n in (1, 2, 3, 5, 17, 64, 150):
        out += [(f"{k}_{n}", v) for k, v in helpers.families(n).items()]
    for _ in range(60):
        n = int(rng.integers(1, 160))
        s = int(rng.choice([1, 2, 3, 4, 256]))
        out.append(("rnd", rng.integers(0, s, size=n, dtype=np.uint8).tobytes()))
    return out


def test_chunked_lyndon_boundaries(oracle):
    for name, x in _cases():
        want = oracle.lyndon_starts(x).tolist()
        for chunk in (1, 4, 16, 64):
            got = model.lyndon_starts_chunked(np.frombuffer(x, np.uint8), chunk).tolist()
            assert got == want, (name, chunk)


def test_doubling_forward(oracle):
    for name, x in _cases():
        assert model.forward(x, chunk=8) == oracle.forward(x), name


def test_doubling_forward_with_tuple_set(oracle):
    """small groups leave the rank-ordered arrays for the text-ordered tuple set (rings by text
    position, rank += number of ring members with a smaller key2)"""
    used = 0
    rng = np.random.default_rng(3)
    cases = _cases()
    for n in (400, 900):
        base = rng.integers(0, 4, size=n // 3, dtype=np.uint8)
        cases.append((f"copies_{n}", bytes(np.concatenate([base, rng.integers(0, 4, size=7, dtype=np.uint8), base, base[: n // 5]]) + 65)))
    for name, x in cases:
        want = oracle.forward(x)
        for tmax in (2, 8, 1000):
            tr = {}
            assert model.forward(x, chunk=8, tmax=tmax, trace=tr) == want, (name, tmax)
            used += tr["entered"]
    assert used > 1000


def test_splitter_inverse(oracle):
    for name, x in _cases():
        for shift in (26, 30, 31):
            assert model.inverse(x, shift=shift) == oracle.inverse(x), (name, shift)


def test_staged_inverse(oracle):
    """the staged single-walk inverse: small slots and warp ranges so that few-byte inputs overflow
    their slots (cont[] + tail walk), flush partial sectors and refill lanes"""
    rng = np.random.default_rng(5)
    cases = _cases()
    for n in (300, 700, 1500):
        cases.append((f"rnd2_{n}", rng.integers(0, 2, size=n, dtype=np.uint8).tobytes()))
        cases.append((f"rnd256_{n}", rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()))
        cases.append((f"fwd_text_{n}", oracle.forward(bytes(rng.integers(97, 101, size=n, dtype=np.uint8)))))
    for name, x in cases:
        want = oracle.inverse(x)
        for shift, slot, Q in ((30, 32, 2), (28, 32, 3), (31, 64, 1), (27, 32, 5), (26, 32, 4), (25, 64, 7), (27, 96, 40), (26, 256, 64)):
            assert model.inverse_staged(x, shift=shift, slot=slot, Q=Q) == want, (name, shift, slot, Q)


def test_ownership_rule_partitions_groups():
    """every group is owned by exactly one worker, and a worker's share is < 2T slots when no
    group has more than T members (k_local_sort_warp: T = 32, k_local_sort_cta: T = 4096)"""
    rng = np.random.default_rng(7)
    for T in (4, 32, 100):
        for _ in range(50):
            sizes = rng.integers(1, T + 1, size=int(rng.integers(1, 60)))
            starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
            gst = np.repeat(starts, sizes)
            ranges = model.owned_ranges(gst, T)
            covered = np.zeros(len(gst), dtype=np.int64)
            for lo, hi in ranges:
                assert hi - lo < 2 * T
                assert lo == len(gst) or gst[lo] == lo          # starts at a group head
                assert hi == len(gst) or gst[hi] == hi          # ends at a group head
                covered[lo:hi] += 1
            assert (covered == 1).all()
    # an oversize group shows up as a share > 2T - 1 somewhere (what k_ls_probe looks for)
    gst = np.zeros(1000, dtype=np.int64)
    assert max(hi - lo for lo, hi in model.owned_ranges(gst, 32)) >= 64


def test_cta_sort_word_orders_by_group_key_slot():
    rng = np.random.default_rng(8)
    for _ in range(20):
        cnt = int(rng.integers(2, 200))
        sizes = []
        while sum(sizes) < cnt:
            sizes.append(int(rng.integers(1, 40)))
        sizes[-1] -= sum(sizes) - cnt
        starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
        g = np.repeat(starts, sizes)
        r = rng.integers(0, 1 << 31, size=cnt)
        r[rng.integers(0, cnt, size=cnt // 3)] = r[0]  # ties
        words = [model.cta_sort_key(g[s], r[s], s) for s in range(cnt)]
        P = 1
        while P < cnt:
            P <<= 1
        got = model.bitonic_sort(words + [(1 << 64) - 1] * (P - cnt))[:cnt]
        want = sorted(range(cnt), key=lambda s: (g[s], r[s], s))
        assert [w & 8191 for w in got] == want
        # groups stay where they were: slot s of the output belongs to the group that owned slot s
        assert [(w >> 44) for w in got] == g.tolist()


def test_binned_scatter_equals_direct_scatter():
    rng = np.random.default_rng(9)
    for n in (256, 1000, 70_001):
        kb = max(8, int(n - 1).bit_length())
        pos = rng.permutation(n).astype(np.int64)
        val = rng.integers(0, 1 << 30, size=n)
        direct = np.zeros(n, dtype=val.dtype)
        direct[pos] = val
        assert np.array_equal(model.binned_scatter(pos, val, n, kb), direct)


def test_counted_bin_scatter_of_a_sparse_set_equals_direct_scatter():
    rng = np.random.default_rng(11)
    for n, m in ((1000, 1), (1000, 300), (70_001, 20_000), (70_001, 70_001)):
        kb = max(8, int(n - 1).bit_length())
        pos = rng.permutation(n)[:m].astype(np.int64)
        val = rng.integers(0, 1 << 30, size=m)
        rank = rng.integers(0, 1 << 30, size=n)
        direct = rank.copy()
        direct[pos] = val
        assert np.array_equal(model.binned_scatter_counted(pos, val, rank, kb), direct)


def test_onesweep_tile_permutation_is_the_stable_partition():
    rng = np.random.default_rng(10)
    for tile in (384 * 12, 100):
        digits = rng.integers(0, 256, size=tile)
        inv = model.onesweep_tile_permutation(digits)
        assert np.array_equal(inv, np.argsort(digits, kind="stable"))


def test_digit_histograms_from_one_window_histogram():
    """the eight digit histograms of the initial sort equal slices of one histogram of the leading symbols
    (k_init_keys + k_digit_hists), for every alphabet width the driver uses it with, with short factors (the
    rotation wraps many times inside the key), narrow keys, and keys whose spare bits hold the top of one more symbol"""
    rng = np.random.default_rng(12)
    used = used_extra = 0
    for bits in (1, 2, 3, 4, 5, 6, 7, 8):
        for keybits in (64, 40, 24, 9):
            k0 = max(1, keybits // bits)
            if k0 * bits > 64:
                continue
            for extra in sorted({0, min(bits - 1, max(0, keybits - k0 * bits))}):
                n = 600
                codes = rng.integers(0, 1 << bits, size=n)
                cuts = sorted(set(rng.integers(1, n, size=12).tolist()) | {1, 2, 5})   # factors of length 1, 1, 3, ...
                starts = [0] + cuts + [n]
                keys = model.initial_keys(codes, starts, bits, k0, extra)
                assert max(keys) < 1 << (k0 * bits + extra)
                got = model.digit_hists_from_windows(keys, bits, k0, extra)
                if got is None:
                    continue
                used += 1
                used_extra += extra > 0
                P0 = -(-(k0 * bits + extra) // 8)
                for p in range(P0):
                    want = np.bincount(np.array([(k >> (8 * p)) & 255 for k in keys]), minlength=256)
                    assert np.array_equal(got[p], want), (bits, k0, extra, p)
    assert used >= 16 and used_extra >= 3, (used, used_extra)


def test_partial_symbol_keys_refine_the_order_consistently():
    """keys with the top bits of symbol k0 + 1 in their spare bits order the rotations consistently with the
    omega-order and tie only where the first k0 symbols are equal: what the doubling rounds need of the initial ranks"""
    rng = np.random.default_rng(13)
    for bits, k0, extra in ((6, 10, 4), (6, 3, 4), (7, 9, 1), (5, 12, 4), (3, 5, 1)):
        n = 300
        codes = rng.integers(0, min(1 << bits, 5), size=n)
        starts = [0, 7, 8, 120, n]
        plain = model.initial_keys(codes, starts, bits, k0, 0)
        fine = model.initial_keys(codes, starts, bits, k0, extra)
        deep = model.initial_keys(codes, starts, bits, k0 + 1, 0)   # one whole symbol more
        for a in range(0, n, 7):
            for b in range(n):
                if fine[a] == fine[b]:
                    assert plain[a] == plain[b]
                if fine[a] < fine[b]:
                    assert deep[a] <= deep[b] and plain[a] <= plain[b]
