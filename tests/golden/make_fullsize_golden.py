"""Full-size golden hashes: SHA-256 of the forward BWTS the UNMODIFIED reference `mk_bwts`
(oracle/_ref, built from /root/reference by oracle/Makefile) produces for the BASELINE
configurations at their full sizes -- C2 64 MiB text, C3 256 MiB tiled text, C4 1 GiB DNA.
Run where /root/reference is mounted (needs ~10 GiB of RAM and ~10 minutes for C4):

    python tests/golden/make_fullsize_golden.py        # rewrites tests/golden/fullsize.json

tests/test_gpu_parity.py compares the CUDA output at the same sizes with these hashes, so the
full-size GPU tests are bit-exact checks against the reference, not only property checks.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import helpers  # noqa: E402

CASES = (("C2", "text", 2, 64 << 20), ("C3", "tiled", 3, 256 << 20), ("C4", "dna", 4, 1 << 30))


def main():
    assert helpers.ref_available(), "oracle/_ref is missing: make -C oracle ref (needs /root/reference)"
    gen = helpers.Generator()
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, kind, seed, n in CASES:
            x = gen.make(kind, seed, n)
            src, dst = os.path.join(td, "in"), os.path.join(td, "out")
            with open(src, "wb") as f:
                f.write(x)
            t = time.time()
            subprocess.check_call([str(helpers.REF_DIR / "mk_bwts"), src, dst])
            dt = time.time() - t
            with open(dst, "rb") as f:
                y = f.read()
            out[name] = {"kind": kind, "seed": seed, "n": n, "input_sha256": hashlib.sha256(x).hexdigest(),
                         "fwd_sha256": hashlib.sha256(y).hexdigest(), "reference_seconds": round(dt, 1)}
            print(name, out[name], flush=True)
    (Path(__file__).parent / "fullsize.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
