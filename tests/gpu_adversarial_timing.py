"""Diagnostic (not a test): time adversarial shapes at 16 / 64 MiB through the GPU path."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
import helpers
bwts = helpers.load_product()
gen = helpers.Generator()
def shapes(n):
    rng = np.random.default_rng(1)
    text = np.frombuffer(gen.make("text", 5, n), dtype=np.uint8).copy()
    z = text.copy(); z[n // 4: n // 4 + n // 8] = 0           # a run of zeros (1/8 of the file) inside text
    yield "zeros_run_in_text", z.tobytes()
    yield "all_a", b"a" * n
    yield "abab", (b"ab" * (n // 2 + 1))[:n]
    yield "descending_runs", bytes(np.repeat(np.arange(255, -1, -1, dtype=np.uint8), n // 256 + 1)[:n])
    yield "random2", bytes(rng.integers(97, 99, size=n, dtype=np.uint8))
    yield "fibonacci", helpers.fibonacci_word(n)
    yield "period_1000", (gen.make("text", 6, 1000) * (n // 1000 + 1))[:n]
with bwts.Context(0) as ctx:
    for n in (16 << 20, 64 << 20):
        for name, x in shapes(n):
            t0 = time.perf_counter(); y = ctx.forward_host(x); t1 = time.perf_counter()
            sf = ctx.stats()
            z = ctx.inverse_host(y); t2 = time.perf_counter()
            si = ctx.stats()
            ok = z == x
            top = sorted(sf["classes"].items(), key=lambda kv: -kv[1]["ms"])[:3]
            print(f"{n>>20:3d} MiB {name:20s} fwd {1e3*(t1-t0):9.1f} ms inv {1e3*(t2-t1):8.1f} ms rt={'ok' if ok else 'BROKEN'} "
                  f"fallback={sf["lyndon_fallback"]} factors={sf["factors"]} rounds={sf["rounds"]} unreached={si['unreached']} "
                  + " ".join(f"{k}={v['ms']:.1f}" for k, v in top), flush=True)
