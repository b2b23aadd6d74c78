"""CPU tests of bench.py's contract: the reference arm prints one JSON line with the agreed keys,
and the GPU arm refuses to run (loudly, non-zero exit) when no CUDA device is present."""
import json
import subprocess
import sys

import pytest

import helpers

BENCH = str(helpers.REPO / "bench.py")


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bwts_round_trip_throughput" and d["unit"] == "MB/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "MB/s"
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, BENCH, "--workload", "C1", "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert not [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")], "no result line without a GPU"
    assert "no CUDA device" in p.stderr or "no CPU path" in p.stderr or "CUDA" in p.stderr
