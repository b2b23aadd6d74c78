"""Diagnostic (not a test): run inputs repeatedly through the GPU path, report mismatches."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
import helpers

bwts = helpers.load_product()
oracle = helpers.Oracle()
gen = helpers.Generator()
cases = [("banana", b"banana"), ("mississippi", b"mississippi"), ("abab100", b"abab" * 100),
         ("text300k", gen.make("text", 2, 300_000)), ("dna100k", gen.make("dna", 4, 100_000)),
         ("text1m", gen.make("text", 7, 1_000_003)), ("rand200k", gen.make("random", 3, 200_000))]
with bwts.Context(0) as ctx:
    for rep in range(3):
        for name, x in cases:
            want = oracle.forward(x)
            got = ctx.forward_host(x)
            st = ctx.stats()
            if got != want:
                a = np.frombuffer(got, np.uint8); b = np.frombuffer(want, np.uint8)
                bad = np.flatnonzero(a != b)
                print(f"FWD MISMATCH rep{rep} {name}: {len(bad)} bytes differ, first at {bad[:8]}, "
                      f"factors={st['factors']} lmax={st['longest_factor']} rounds={st['rounds']} "
                      f"bits={st['alphabet_bits']} k0={st['initial_depth']}", flush=True)
                want_f = len(oracle.lyndon_starts(x))
                print(f"   oracle factors={want_f}; hist equal={np.array_equal(np.bincount(a,minlength=256), np.bincount(b,minlength=256))}")
            else:
                print(f"fwd ok rep{rep} {name} factors={st['factors']} rounds={st['rounds']} ms={st['total_ms']:.3f}", flush=True)
            wi = oracle.inverse(x)
            gi = ctx.inverse_host(x)
            st = ctx.stats()
            if gi != wi:
                a = np.frombuffer(gi, np.uint8); b = np.frombuffer(wi, np.uint8)
                bad = np.flatnonzero(a != b)
                print(f"INV MISMATCH rep{rep} {name}: {len(bad)} differ first {bad[:8]} cycles={st['factors']} spl={st['splitters']} unreached={st['unreached']}", flush=True)
            else:
                print(f"inv ok rep{rep} {name} cycles={st['factors']} spl={st['splitters']} unreached={st['unreached']} ms={st['total_ms']:.3f}", flush=True)
