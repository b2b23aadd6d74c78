#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2b; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 120 python tests/gpu_experiments.py C4 base 12:1 > $out/exp_c4.txt 2>&1
timeout 90 python tests/gpu_experiments.py C2 base 12:1 > $out/exp_c2.txt 2>&1
BWTS_B200_TRACE=1 timeout 120 python tests/gpu_experiments.py C3 base > $out/exp_c3.txt 2> $out/exp_c3_trace.txt
timeout 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" >> $out/bench_default.err
tail -3 $out/pytest.txt; tail -3 $out/bench_default.err
