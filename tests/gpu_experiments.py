"""A/B runs of bwts_b200_tune settings on one input (GPU box only; not a test, not the bench).

    python tests/gpu_experiments.py C4 base 10:32 10:128 11:1 ...   > gpurun_out/exp.txt

Every argument after the workload is one experiment: `base` or a comma-separated list of
key:value pairs for bwts_b200_tune (all keys are reset to 0 between experiments).  Prints the
forward / inverse device times and the per-class milliseconds of the second of two runs.
Experiments whose name ends in `!` skip the round-trip check (timing-only settings).
"""
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))

import torch  # noqa: E402

import bench  # noqa: E402
import helpers  # noqa: E402

KEYS = range(0, 32)


def main():
    wl = sys.argv[1]
    exps = sys.argv[2:] or ["base"]
    kind, seed, n, desc = bench.WORKLOADS[wl]
    bwts = helpers.load_product()
    dev = torch.device("cuda", 0)
    t = time.time()
    data = helpers.Generator().make(kind, seed, n)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    d_mid = torch.empty_like(d_in)
    d_back = torch.empty_like(d_in)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"# {desc}: generated in {time.time() - t:.1f} s", flush=True)
    ctx = bwts.Context(0)
    ctx.reserve(n)
    sh = torch.cuda.current_stream(dev).cuda_stream
    ref_mid = None
    for exp in exps:
        check = not exp.endswith("!")
        spec = exp.rstrip("!")
        for k in KEYS:
            try:
                bwts.tune(k, 0)
            except Exception:
                pass
        if spec != "base":
            for kv in spec.split(","):
                k, v = kv.split(":")
                bwts.tune(int(k), int(v))
        for rep in range(2):
            flush.zero_()
            ctx.forward_device(d_in.data_ptr(), n, d_mid.data_ptr(), sh)
            sf = ctx.stats()
            flush.zero_()
            ctx.inverse_device(d_mid.data_ptr(), n, d_back.data_ptr(), sh)
            si = ctx.stats()
        torch.cuda.synchronize(dev)
        ok = ""
        if check:
            ok = "roundtrip=" + str(bool(torch.equal(d_back, d_in)))
            if ref_mid is None:
                ref_mid = d_mid.clone()
            else:
                ok += " fwd_same_as_first=" + str(bool(torch.equal(d_mid, ref_mid)))
        print(f"== {wl} {exp}: fwd {sf['total_ms']:.2f} ms  inv {si['total_ms']:.2f} ms  launches {sf['launches']}+{si['launches']}  "
              f"rounds {sf['rounds']} {ok}", flush=True)
        for name, st in (("fwd", sf), ("inv", si)):
            for cname, c in sorted(st["classes"].items(), key=lambda kv: -kv[1]["ms"]):
                gbs = c["bytes"] / c["ms"] * 1e-6 if c["ms"] > 0 else 0
                print(f"   {name} {cname:14s} {c['launches']:4d} launches {c['ms']:9.3f} ms {gbs:8.1f} GB/s")
    ctx.close()


if __name__ == "__main__":
    main()
