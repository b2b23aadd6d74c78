#!/usr/bin/env python
"""bench.py -- BWTS forward + inverse throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C4|C1]
    (N > 1: launched by torch.distributed.run, one rank per GPU)

A step = one pass of the hot path over one block: forward BWTS of the block, then the
inverse BWTS of the result, both with buffers resident in HBM (`value`), and the same
through the host-buffer C-ABI call with pinned host memory and both copies timed (`e2e`).
Every rank owns one independent block (weak scaling, no collective on the data path);
value = bytes all ranks processed / max-over-ranks device time.

`--impl reference` times the reference's own CPU tools (oracle/_ref, built from the
unmodified sources; else the oracle port) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "tests"))

MB = 1e6
WORKLOADS = {
    # name: (generator kind, seed, bytes, description)                      BASELINE.json configs[i]
    "C1": ("random", 1, 1 << 20, "C1: 1 MiB uniform random bytes"),
    "C2": ("text", 2, 64 << 20, "C2: 64 MiB English-like order-2 Markov text"),
    "C3": ("tiled", 3, 256 << 20, "C3: 256 MiB 64 KiB-tiled text, one substitution per MiB"),
    "C4": ("dna", 4, 1 << 30, "C4: 1 GiB DNA (ACGT) with copied segments"),
    "C3F": ("fibonacci", 0, 256 << 20, "C3 stress: 256 MiB Fibonacci word"),
    # C5: a multi-block file of 8 x 256 MiB independent blocks (seeds 50..57) dealt over the ranks
    "C5": ("text", 50, 256 << 20, "C5: multi-block file, 8 x 256 MiB order-2 Markov text blocks"),
    # the same shape at an eighth of the size (pipeline diagnostics)
    "C5S": ("text", 50, 32 << 20, "C5 small: multi-block file, 8 x 32 MiB order-2 Markov text blocks"),
}
C5_BLOCKS = 8


def block_owner(block, ndev):
    """Round-robin dealing of independent blocks, the same rule as bwts_b200_*_blocks (run_blocks)."""
    return block % ndev


def plan_blocks(workload, rank, world):
    """(generator kind, seed, bytes) of every block this rank transforms in one step.
    C1..C4: one block per rank (weak scaling).  C5: 8 fixed blocks dealt round-robin (strong)."""
    kind, seed, n, _ = WORKLOADS[workload]
    if workload in ("C5", "C5S"):
        return [(kind, seed + b, n) for b in range(C5_BLOCKS) if block_owner(b, world) == rank]
    return [(kind, seed + 100 * rank, n)]


def total_blocks(workload, world):
    return C5_BLOCKS if workload in ("C5", "C5S") else world
DOMINANT = "onesweep_pass"


_JSON_OUT = None


def claim_stdout():
    """Keep the real stdout for the one JSON line; whatever libraries print to fd 1 (NCCL's
    version banner, for one) goes to stderr instead."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit_json(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_product():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bwts_b200", REPO / "bijective-bwt_b200" / "bwts_b200.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bwts_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_input(kind, seed, n):
    import helpers
    return helpers.Generator().make(kind, seed, n)


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from before the warm-up on;
    stop() keeps the samples whose timestamps fall inside the timed regions."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0 = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """the timed region starts now"""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                except Exception:
                    continue
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            sm.sort()
            return sm, mx, reasons

        inside = [x for x in self.rows if self.t0 is not None and self.t0 - 0.05 <= x[0] <= t1 + 0.1]
        sm, mx, reasons = digest(inside)
        window = "timed regions"
        if not sm:  # region shorter than the sampling period: report what was seen around it
            sm, mx, reasons = digest(self.rows)
            window = "whole run (timed region shorter than the sampling period)"
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arms

def ref_kind():
    import helpers
    return "reference" if helpers.ref_available() else "port"


def cpu_round_trip(data, kind):
    """forward + inverse of `data` on one host core; returns seconds (fwd, inv) and checks the round trip."""
    import helpers
    if kind == "reference":
        d = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=d) as td:
            src, mid, back = Path(td) / "in", Path(td) / "mid", Path(td) / "back"
            src.write_bytes(data)
            t0 = time.perf_counter()
            subprocess.check_call([str(helpers.REF_DIR / "mk_bwts"), str(src), str(mid)])
            t1 = time.perf_counter()
            subprocess.check_call([str(helpers.REF_DIR / "unbwts"), str(mid), str(back)])
            t2 = time.perf_counter()
            assert back.read_bytes() == data, "reference round trip failed"
            fwd = mid.read_bytes()
    else:
        o = helpers.Oracle()
        t0 = time.perf_counter()
        fwd = o.forward(data)
        t1 = time.perf_counter()
        back = o.inverse(fwd)
        t2 = time.perf_counter()
        assert back == data, "oracle round trip failed"
    return t1 - t0, t2 - t1, fwd


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    kind_name, seed, n, desc = WORKLOADS[args.workload]
    kind = ref_kind()
    total = max(1, args.steps + args.warmup)
    # ~2.5 MB/s round trip on one core; keep the whole run near two minutes
    budget = int(2.5e6 * 110 / total)
    sample = 1 << 20
    while sample * 2 <= min(n, budget):
        sample *= 2
    nproc = os.cpu_count() or 1
    # the same blocks the GPU arm transforms in one step, run side by side on the host cores
    all_blocks = [blk for r in range(world) for blk in plan_blocks(args.workload, r, world)]
    datas = [make_input(k_, s_, n_)[:sample] for k_, s_, n_ in all_blocks]
    blocks = len(datas)
    conc = min(blocks, nproc)

    def one_step():
        t0 = time.perf_counter()
        if conc == 1:
            for d in datas:
                cpu_round_trip(d, kind)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(conc) as ex:
                list(ex.map(lambda d: cpu_round_trip(d, kind), datas))
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = blocks * sample / MB / sec
    sample_desc = (f"first {sample >> 20} MiB of each of {blocks} block(s) of {desc}; "
                   f"{'unmodified reference mk_bwts+unbwts (oracle/_ref; suffix sort = substitute SA-IS, not libdivsufsort)' if kind == 'reference' else 'oracle port'}; "
                   f"{conc} process(es) at a time, 1 thread each")
    line = {
        "impl": "reference", "metric": "bwts_round_trip_throughput", "value": value, "unit": "MB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.workload in ("C5", "C5S") else "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc + ", forward + inverse round trip", "bytes_per_gpu": n * blocks // world,
                   "blocks": blocks, "sample_bytes": sample},
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": conc, "kind": kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# --------------------------------------------------------------------------- GPU arm

def run_gpu_arm(args, rank, local_rank, world):
    import torch
    bwts = load_product()
    if bwts.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; the product has no CPU path")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    kind_name, seed, n, desc = WORKLOADS[args.workload]
    plan = plan_blocks(args.workload, rank, world)
    if args.bytes:
        n = args.bytes
        plan = [(k_, s_, n) for k_, s_, _ in plan]
    for kv in args.tune:
        k, v = kv.split(":")
        bwts.tune(int(k), int(v))
    ctx = bwts.Context(local_rank)
    ctx.reserve(n)
    blocks = []
    for k_, s_, n_ in plan:
        data = make_input(k_, s_, n_)
        host_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
        blocks.append({"data": data, "host_in": host_in, "host_mid": torch.empty(n_, dtype=torch.uint8).pin_memory(),
                       "host_back": torch.empty(n_, dtype=torch.uint8).pin_memory(), "d_in": host_in.to(dev),
                       "d_mid": torch.empty(n_, dtype=torch.uint8, device=dev),
                       "d_back": torch.empty(n_, dtype=torch.uint8, device=dev), "n": n_})
    my_bytes = sum(b["n"] for b in blocks)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    sh = stream.cuda_stream

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    agg = {}

    def absorb(st, acc):
        for name, c in st["classes"].items():
            a = acc.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0.0})
            a["launches"] += c["launches"]; a["ms"] += c["ms"]; a["bytes"] += c["bytes"]

    def device_step(acc=None):
        """forward + inverse of every block this rank owns; returns (fwd_ms events, inv_ms events, launches)"""
        out = []
        for b in blocks:
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e2 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.forward_device(b["d_in"].data_ptr(), b["n"], b["d_mid"].data_ptr(), sh)
            sf = ctx.stats()
            e1.record(stream)
            ctx.inverse_device(b["d_mid"].data_ptr(), b["n"], b["d_back"].data_ptr(), sh)
            si = ctx.stats()
            e2.record(stream)
            if acc is not None:
                absorb(sf, acc); absorb(si, acc)
            out.append((e0, e1, e2, sf, si))
        return out

    def check_round_trip():
        for b in blocks:
            assert torch.equal(b["d_back"], b["d_in"]), "round trip lost data"

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up (also the correctness gate of the bench itself)
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize(dev)
    if args.warmup:
        check_round_trip()

    # ---- timed region: exactly K steps, device resident
    barrier()
    sampler.mark()
    t_begin = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    evs = []
    launches = 0
    last_f = last_i = None
    for _ in range(args.steps):
        for e0, e1, e2, sf, si in device_step(agg):
            launches += sf["launches"] + si["launches"]
            evs.append((e0, e1, e2))
            last_f, last_i = sf, si
    t_end.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    total_ms = t_begin.elapsed_time(t_end)
    fwd_ms = sum(a.elapsed_time(b) for a, b, _ in evs) / args.steps
    inv_ms = sum(b.elapsed_time(c) for _, b, c in evs) / args.steps
    check_round_trip()

    # ---- end to end: the host-buffer C-ABI call, pinned host memory, both copies inside
    def e2e_step():
        t0 = time.perf_counter()
        dev_ms = 0.0
        for b in blocks:
            ctx.forward_host_ptr(b["host_in"].data_ptr(), b["n"], b["host_mid"].data_ptr())
            sf = ctx.stats()
            ctx.inverse_host_ptr(b["host_mid"].data_ptr(), b["n"], b["host_back"].data_ptr())
            si = ctx.stats()
            dev_ms += sum(s["h2d_ms"] + s["total_ms"] + s["d2h_ms"] for s in (sf, si))
        return (time.perf_counter() - t0) * 1e3, dev_ms

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    e2e = [e2e_step() for _ in range(args.steps)]
    e2e_wall_ms = sum(w for w, _ in e2e) / args.steps
    e2e_dev_ms = sum(d for _, d in e2e) / args.steps
    for b in blocks:
        assert torch.equal(b["host_back"], b["host_in"]), "e2e round trip lost data"

    # ---- several blocks per GPU: the block pipeline (bwts_b200_*_blocks: H2D of block b+1 |
    # transform of block b | D2H of block b-1), whole call timed on the host clock
    pipe_ms = None
    if len(blocks) > 1 and len({b["n"] for b in blocks}) == 1:
        nb_, bl = len(blocks), blocks[0]["n"]
        cat_in = torch.cat([b["host_in"] for b in blocks]).pin_memory()
        cat_mid = torch.empty_like(cat_in).pin_memory()
        cat_back = torch.empty_like(cat_in).pin_memory()

        def pipe_step():
            t0 = time.perf_counter()
            bwts.blocks_ptr(0, cat_in.data_ptr(), nb_ * bl, bl, cat_mid.data_ptr(), devices=[local_rank])
            bwts.blocks_ptr(1, cat_mid.data_ptr(), nb_ * bl, bl, cat_back.data_ptr(), devices=[local_rank])
            return (time.perf_counter() - t0) * 1e3

        for _ in range(min(args.warmup, 1)):
            pipe_step()
        barrier()
        pipe_ms = sum(pipe_step() for _ in range(args.steps)) / args.steps
        assert torch.equal(cat_back, cat_in), "pipelined e2e round trip lost data"
        for i, b in enumerate(blocks):
            assert torch.equal(cat_mid[i * bl:(i + 1) * bl], b["host_mid"]), "pipelined forward differs from the per-block call"
        del cat_in, cat_mid, cat_back
    clocks = sampler.stop()  # sampled across the timed regions (device-resident and end-to-end)
    barrier()
    data, d_mid = blocks[0]["data"], blocks[0]["d_mid"]

    # ---- max over ranks (times), sum over ranks (bytes)
    t = torch.tensor([total_ms, fwd_ms, inv_ms, e2e_wall_ms, e2e_dev_ms, pipe_ms or 0.0], dtype=torch.float64, device=dev)
    nb = torch.tensor([float(my_bytes)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nb, op=dist.ReduceOp.SUM)
    total_ms, fwd_ms, inv_ms, e2e_wall_ms, e2e_dev_ms, pipe_ms_max = t.tolist()
    job_bytes = nb.item()
    ms_per_step = total_ms / args.steps

    if rank == 0:
        peak, peak_src = measured_peak()
        dom = agg.get(DOMINANT, {"launches": 0, "ms": 0.0, "bytes": 0.0})
        achieved = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
        traffic = None
        tp = REPO / "profiles" / "onesweep_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "bwts_round_trip_throughput", "value": job_bytes / MB / (ms_per_step * 1e-3), "unit": "MB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if args.workload in ("C5", "C5S") else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc + ", forward + inverse round trip", "bytes_per_gpu": my_bytes,
                       "blocks": total_blocks(args.workload, world), "l2": "flushed between steps (256 MiB device memset inside the timed loop)",
                       "generator": f"bijective-bwt_b200/host/gen_input.c kind={kind_name} seed={seed}"
                                    + ("+block" if args.workload in ("C5", "C5S") else "+100*rank"),
                       "parallelism": f"independent blocks over {world} GPU(s), round-robin, no collective"},
            "forward_mbs": job_bytes / MB / (fwd_ms * 1e-3), "inverse_mbs": job_bytes / MB / (inv_ms * 1e-3),
            "forward_ms": fwd_ms, "inverse_ms": inv_ms,
            "e2e": {"value": job_bytes / MB / (e2e_dev_ms * 1e-3), "unit": "MB/s",
                    "h2d_bytes_per_step": 2 * my_bytes, "d2h_bytes_per_step": 2 * my_bytes,
                    "wall_value": job_bytes / MB / (e2e_wall_ms * 1e-3),
                    "how": "bwts_b200_forward_host + bwts_b200_inverse_host on pinned host buffers; CUDA events "
                           "from before the H2D copy to after the D2H copy (wall_value: host clock)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_onesweep_pass", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": peak_src, "launches": dom["launches"],
                         "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"] if dom["launches"] else 0,
                         "avg_launch_ms": dom["ms"] / dom["launches"] if dom["launches"] else 0,
                         "share_of_step": dom["ms"] / (total_ms) if total_ms else None},
            "kernel_classes": {k: {"launches": v["launches"] // max(args.steps, 1), "ms_per_step": v["ms"] / args.steps,
                                   "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
                               for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
            "transform": {"factors": last_f["factors"], "longest_factor": last_f["longest_factor"],
                          "alphabet_bits": last_f["alphabet_bits"], "initial_depth": last_f["initial_depth"],
                          "doubling_rounds": last_f["rounds"], "radix_passes": last_f["radix_passes"],
                          "local_sort_rounds": last_f["local_rounds"], "cta_sort_rounds": last_f["cta_rounds"],
                          "live_sum": last_f["live_sum"], "cycles": last_i["factors"],
                          "splitters": last_i["splitters"], "unreached": last_i["unreached"]},
        }
        if pipe_ms is not None:
            # the headline end-to-end number of a multi-block workload is the pipelined call
            line["e2e"]["one_block_at_a_time_value"] = line["e2e"]["value"]
            line["e2e"]["value"] = job_bytes / MB / (pipe_ms_max * 1e-3)
            line["e2e"]["wall_value"] = line["e2e"]["value"]
            line["e2e"]["how"] = ("bwts_b200_forward_blocks + bwts_b200_inverse_blocks over all blocks of the rank, "
                                  "pinned host buffers, per-device pipeline H2D | transform | D2H; host clock around "
                                  "both calls (one_block_at_a_time_value: the per-block host-buffer calls, CUDA events)")
        if world == 1 and not args.no_cpu:
            kind = ref_kind()
            sample = min(n, args.cpu_sample)
            tf, ti, fwd = cpu_round_trip(data[:sample], kind)
            # the sample doubles as a full-size parity check of the timed workload
            if sample == n:
                assert bytes(d_mid.cpu().numpy()) == fwd, "GPU forward differs from the CPU baseline output"
            line["cpu_baseline"] = {
                "value": sample / MB / (tf + ti), "unit": "MB/s", "cores": 1, "kind": kind,
                "forward_mbs": sample / MB / tf, "inverse_mbs": sample / MB / ti,
                "sample": f"first {sample} bytes of the same block, forward + inverse, "
                          + ("unmodified reference mk_bwts + unbwts from oracle/_ref (suffix sort = substitute SA-IS, "
                             "not libdivsufsort), 1 thread" if kind == "reference" else "oracle port, 1 thread"),
                "host_cpus": os.cpu_count()}
        emit_json(line)
    ctx.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--bytes", type=int, default=0, help="override the block size (diagnostics only)")
    ap.add_argument("--cpu-sample", type=int, default=64 << 20, help="bytes of the block the CPU baseline runs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tune", action="append", default=[], help="key:value for bwts_b200_tune (experiments)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        claim_stdout()
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # not under torchrun: re-launch ourselves with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    claim_stdout()
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
