"""Randomised differential stress run (not collected by pytest): GPU path vs oracle.
usage: python tests/gpu_stress.py [seconds] [seed]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
import helpers

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bwts = helpers.load_product()
oracle = helpers.Oracle()
gen = helpers.Generator()


def make_case():
    kind = rng.integers(0, 9)
    n = int(rng.choice([rng.integers(1, 300), rng.integers(300, 20_000), rng.integers(20_000, 400_000)]))
    if kind == 0:
        sigma = int(rng.choice([1, 2, 3, 4, 5, 16, 64, 256]))
        return f"iid{sigma}", rng.integers(0, sigma, size=n, dtype=np.uint8).tobytes()
    if kind == 1:
        return "text", gen.make("text", int(rng.integers(1, 1 << 30)), n)
    if kind == 2:
        return "dna", gen.make("dna", int(rng.integers(1, 1 << 30)), n)
    if kind == 3:
        return "tiled", gen.make("tiled", int(rng.integers(1, 1 << 30)), n)
    if kind == 4:  # short period with a few mutations
        p = int(rng.integers(1, 50))
        base = rng.integers(97, 100, size=p, dtype=np.uint8)
        x = np.resize(base, n).copy()
        for _ in range(int(rng.integers(0, 4))):
            x[int(rng.integers(0, n))] = int(rng.integers(0, 256))
        return f"period{p}", x.tobytes()
    if kind == 5:  # runs
        runs = rng.integers(0, 4, size=n // 5 + 1, dtype=np.uint8)
        lens = rng.integers(1, 40, size=len(runs))
        return "runs", np.repeat(runs, lens)[:n].tobytes() or b"a"
    if kind == 6:  # repeated blocks w w' w
        w = rng.integers(0, 3, size=max(1, n // 3), dtype=np.uint8).tobytes()
        return "ww", (w + w[: len(w) // 2] + w)[:n] or b"a"
    if kind == 7:
        return "fib", helpers.fibonacci_word(n)
    return "descending", bytes(sorted(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes(), reverse=True))


t_end = time.time() + budget
cases = bad = 0
with bwts.Context(0) as ctx:
    while time.time() < t_end:
        name, x = make_case()
        if not x:
            continue
        chunk = int(rng.choice([0, 0, 1, 3, 16, 100, 512, 4096]))
        shift = int(rng.choice([0, 0, 26, 28, 30, 31]))
        bwts.tune(0, chunk); bwts.tune(1, shift)
        bwts.tune(3, int(rng.random() < 0.15)); bwts.tune(4, int(rng.random() < 0.15))
        # round-2 paths: binned scatter, CTA sort on/off + kind, emit form, inverse form + marks, tuple set size + form,
        # Lyndon scan
        knobs = {7: int(rng.choice([0, 0, 1, 2, 4, 4])), 8: int(rng.random() < 0.2), 9: int(rng.choice([0, 1, 2, 3])),
                 12: int(rng.random() < 0.2), 14: int(rng.choice([0, 1, 2, 3, 8, 32])), 15: int(rng.choice([0, 1, 2])),
                 17: int(rng.choice([0, 1, 2])), 18: int(rng.random() < 0.3), 20: int(rng.random() < 0.4),
                 21: int(rng.random() < 0.3), 22: int(rng.random() < 0.3), 6: int(rng.choice([0, 0, 0, 24, 40, 56]))}
        for key, val in knobs.items():
            bwts.tune(key, val)
        wf, wi = oracle.forward(x), oracle.inverse(x)
        gf, gi = ctx.forward_host(x), ctx.inverse_host(x)
        cases += 1
        sa_ok = True
        if cases % 7 == 0 and len(x) > 1:
            sa_ok = bool(np.array_equal(bwts.suffix_array(x), oracle.suffix_array(x)))
        if gf != wf or gi != wi or not sa_ok or ctx.inverse_host(gf) != x:
            bad += 1
            Path("gpurun_out").mkdir(exist_ok=True)
            Path(f"gpurun_out/stress_fail_{bad}.bin").write_bytes(x)
            print(f"MISMATCH {name} n={len(x)} chunk={chunk} shift={shift} knobs={knobs} fwd_ok={gf == wf} inv_ok={gi == wi} sa_ok={sa_ok}", flush=True)
            if bad >= 5:
                break
print(f"stress: {cases} cases, {bad} mismatches, seed {seed}")
sys.exit(1 if bad else 0)
