"""CPU test of the N > 1 path's host logic with a real process group (gloo, world_size 2):
block dealing (the rule bench.py and bwts_b200_*_blocks share), no data-path collective, the
max-over-ranks / sum-over-ranks reductions of the bench, and that per-rank outputs put back
at their block offsets equal the block-wise transform.  The per-block transform here is the
ORACLE standing in for the GPU (this is a test of the plumbing, not of the product path)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))


def _worker(rank, world, port, nblocks, block_len, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import helpers
    oracle = helpers.Oracle()
    gen = helpers.Generator()
    whole = gen.make("text", 99, nblocks * block_len - 1234)   # last block is ragged
    mine = [b for b in range(nblocks) if bench.block_owner(b, world) == rank]
    out = {}
    for b in mine:
        blk = whole[b * block_len:(b + 1) * block_len]
        out[b] = oracle.forward(blk)
    # bench-style reductions: time = max over ranks, bytes = sum over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    nb = torch.tensor([float(sum(len(v) for v in out.values()))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(nb, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        merged = {}
        for d in gathered:
            assert not (set(d) & set(merged)), "a block was dealt to two ranks"
            merged.update(d)
        assert sorted(merged) == list(range(nblocks))
        got = b"".join(merged[b] for b in range(nblocks))
        want = b"".join(oracle.forward(whole[o:o + block_len]) for o in range(0, len(whole), block_len))
        q.put((got == want, t.item(), nb.item(), len(whole)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_deal_blocks_and_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, nblocks, block_len = 2, 5, 40_000
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, nblocks, block_len, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, tmax, nbytes, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert tmax == 11.0            # max over ranks, not rank 0's own time
    assert nbytes == float(total)  # bytes of all ranks


def test_block_plans_cover_every_block_once():
    import bench
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += [s for _, s, _, _ in bench.plan_blocks("C5", r, world)]
        assert sorted(seen) == list(range(50, 58))
        assert len({s for r in range(world) for _, s, _, _ in bench.plan_blocks("C2", r, world)}) == world
        assert bench.default_workload(world) == ("C4" if world == 1 else "C5")
