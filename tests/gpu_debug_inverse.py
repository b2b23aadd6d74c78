"""GPU box only: which input / setting makes the inverse fail (run with BWTS_B200_SYNC=1)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers  # noqa: E402

bwts = helpers.load_product()
oracle = helpers.Oracle()
gen = helpers.Generator()
rng = np.random.default_rng(13)
cases = [b"a" * 100_000, b"ab" * 60_000, bytes(rng.integers(0, 2, size=200_000, dtype=np.uint8)),
         gen.make("text", 91, 1_000_000), gen.make("dna", 92, 1_500_000), helpers.fibonacci_word(120_000),
         bytes(np.repeat(rng.integers(0, 256, size=3000, dtype=np.uint8), 40))]
for ci, x in enumerate(cases):
    want = oracle.inverse(x)
    for mark in (0, 1):
        for shift in (0, 22, 29):
            ctx = bwts.Context(0)
            bwts.tune(15, mark)
            bwts.tune(1, shift)
            try:
                got = ctx.inverse_host(x)
                st = ctx.stats()
                print(ci, len(x), mark, shift, "ok" if got == want else "MISMATCH", st["unreached"], st["splitters"], st["inverse_attempts"], flush=True)
            except Exception as e:
                print(ci, len(x), mark, shift, "ERROR", e, ctx.last_cuda_error(), flush=True)
            try:
                ctx.close()
            except Exception:
                pass
