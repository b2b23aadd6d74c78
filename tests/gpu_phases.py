"""Diagnostic (not a test): per-phase cycle counters of the re-rank and onesweep kernels over one
forward transform.  Needs the instrumented build:  make -C bijective-bwt_b200 profile-lib
    BWTS_B200_LIB=bijective-bwt_b200/libbwts_b200_ph.so python tests/gpu_phases.py C2
"""
import ctypes
import sys

import helpers

sys.path.insert(0, str(helpers.REPO))
import bench  # noqa: E402

bwts = helpers.load_product()
kind, seed, n, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
x = helpers.Generator().make(kind, seed, n)
L = bwts.lib()
ph = (ctypes.c_ulonglong * 32)()
with bwts.Context(0) as ctx:
    ctx.forward_host(x)
    L.bwts_b200_debug_phases(ph, 1)
    ctx.forward_host(x)
    st = ctx.stats()
    L.bwts_b200_debug_phases(ph, 1)
names = {0: "os zero", 1: "os loads", 2: "os rank", 3: "os sync", 4: "os digit-scan", 5: "os stage", 6: "os look-back",
         7: "os sync", 8: "os write", 16: "rr loads+flags", 17: "rr sync", 18: "rr keep/route", 19: "rr scans",
         20: "rr look-back", 21: "rr sync", 22: "rr ranks+staging", 23: "rr sync", 24: "rr stream writes"}
print(desc, "forward", st["total_ms"], "ms")
for lo, hi, label in ((0, 9, "onesweep"), (16, 25, "re-rank")):
    tot = sum(ph[i] for i in range(lo, hi))
    print(label, "total Mcycles", tot / 1e6)
    for i in range(lo, hi):
        print("   %-18s %6.1f%%" % (names[i], 100.0 * ph[i] / max(tot, 1)))
