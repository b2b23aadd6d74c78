"""Full-size golden hashes: SHA-256 of the forward BWTS the UNMODIFIED reference `mk_bwts`
(oracle/_ref, built from /root/reference by oracle/Makefile) produces for the BASELINE
configurations at their full sizes -- C2 64 MiB text, C3 256 MiB tiled text, the 256 MiB
Fibonacci word SURVEY 8(d) names as the C3 stress variant, C4 1 GiB DNA, the eight 256 MiB
blocks of the C5 multi-block file (seeds 50..57), and C6, a 1.5 GiB DNA file above the 2^30
limit of round 1, and C7, a DNA file of 2^31 - 1 bytes, the largest the reference accepts (len < 2^31).
Run where /root/reference is mounted (C4 needs ~10 GiB of RAM and ~7 minutes, C6 ~15 GiB):

    python tests/golden/make_fullsize_golden.py [--only C5_0,C5_1,...] [--jobs 4]

Existing entries of tests/golden/fullsize.json are kept unless named in --only (or the file is
absent).  tests/test_gpu_parity.py and bench.py's warm-up compare the CUDA output at the same
sizes with these hashes, so the full-size GPU runs are bit-exact checks against the reference,
not only property checks.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import helpers  # noqa: E402

CASES = [("C2", "text", 2, 64 << 20), ("C3", "tiled", 3, 256 << 20), ("C3F", "fibonacci", 0, 256 << 20),
         ("C4", "dna", 4, 1 << 30)]
CASES += [(f"C5_{b}", "text", 50 + b, 256 << 20) for b in range(8)]
CASES += [("C6", "dna", 6, 3 << 29), ("C7", "dna", 7, (1 << 31) - 1)]
OUT = Path(__file__).parent / "fullsize.json"


def one(case):
    name, kind, seed, n = case
    gen = helpers.Generator()
    x = gen.make(kind, seed, n)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=d) as td:
        src, dst = os.path.join(td, "in"), os.path.join(td, "out")
        with open(src, "wb") as f:
            f.write(x)
        t = time.time()
        subprocess.check_call([str(helpers.REF_DIR / "mk_bwts"), src, dst])
        dt = time.time() - t
        h = hashlib.sha256()
        with open(dst, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
    rec = {"kind": kind, "seed": seed, "n": n, "input_sha256": hashlib.sha256(x).hexdigest(),
           "fwd_sha256": h.hexdigest(), "reference_seconds": round(dt, 1)}
    print(name, rec, flush=True)
    return name, rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated case names (default: the missing ones)")
    ap.add_argument("--jobs", type=int, default=1)
    args = ap.parse_args()
    assert helpers.ref_available(), "oracle/_ref is missing: make -C oracle ref (needs /root/reference)"
    out = json.loads(OUT.read_text()) if OUT.exists() else {}
    want = set(args.only.split(",")) if args.only else {c[0] for c in CASES if c[0] not in out}
    todo = [c for c in CASES if c[0] in want]
    with ThreadPoolExecutor(max(1, args.jobs)) as ex:
        for name, rec in ex.map(one, todo):
            out[name] = rec
            OUT.write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
