"""Diagnostic (CPU, not a test): how the rotations of the DNA workload split into the three live sets after the
initial sort.  Counts the multiplicities of the 32-mers (= the initial 64-bit key at 2 bits per symbol; factor wrap
ignored) of the first 2^lg bytes of the C4 generator.

    python tests/diag_group_sizes.py 28
    elements in groups > 1 : 75.6 %   > 8 : 33.6 %   > 32 : 14.9 %   > 4096 : 0.50 %   > 8192 : 0.48 % (52 groups)

i.e. tuple set 42 %, S set 18 %, L set 15 % of the positions; the oversize groups that keep the L set off the
CTA-local sort in rounds 1-3 hold 3 % of it.
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import helpers  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    n = 1 << lg
    x = np.frombuffer(helpers.Generator().make("dna", 4, n), dtype=np.uint8)
    code = np.zeros(256, np.uint64)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    cc = np.concatenate([code[x], np.zeros(32, np.uint64)])
    key = np.zeros(n, np.uint64)
    for j in range(32):
        key = (key << np.uint64(2)) | cc[j:j + n]
    t = time.time()
    key.sort()
    print("sorted in %.1f s" % (time.time() - t))
    b = np.flatnonzero(np.concatenate([[True], key[1:] != key[:-1], [True]]))
    sz = np.diff(b)
    for thr in (1, 8, 32, 4096, 8192, 65536):
        big = sz[sz > thr]
        print("elements in groups > %5d: %11d  %7.3f %%   groups %d" % (thr, int(big.sum()), 100.0 * big.sum() / n, len(big)))
    print("largest group", int(sz.max()))


if __name__ == "__main__":
    main()
