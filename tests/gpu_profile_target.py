"""GPU box only: one forward + one inverse of a BASELINE workload, nothing else -- the command ncu wraps.

    python tests/gpu_profile_target.py C4 [reps]
"""
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))

import torch  # noqa: E402

import bench  # noqa: E402
import helpers  # noqa: E402


def main():
    wl = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    kind, seed, n, desc = bench.WORKLOADS[wl]
    bwts = helpers.load_product()
    dev = torch.device("cuda", 0)
    data = helpers.Generator().make(kind, seed, n)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    d_mid = torch.empty_like(d_in)
    d_back = torch.empty_like(d_in)
    ctx = bwts.Context(0)
    ctx.reserve(n)
    sh = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(reps):
        ctx.forward_device(d_in.data_ptr(), n, d_mid.data_ptr(), sh)
        sf = ctx.stats()
        ctx.inverse_device(d_mid.data_ptr(), n, d_back.data_ptr(), sh)
        si = ctx.stats()
    torch.cuda.synchronize(dev)
    assert torch.equal(d_back, d_in), "round trip lost data"
    print(f"{desc}: fwd {sf['total_ms']:.2f} ms ({sf['launches']} launches), inv {si['total_ms']:.2f} ms ({si['launches']} launches)")
    ctx.close()


if __name__ == "__main__":
    main()
