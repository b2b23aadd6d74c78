#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2f; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 100 python tests/gpu_experiments.py C4 base 14:1 > $out/exp_c4.txt 2>&1
timeout 100 python tests/gpu_experiments.py C5 base 14:1 > $out/exp_c5.txt 2>&1
timeout 100 python tests/gpu_experiments.py C3 base > $out/exp_c3.txt 2>&1
timeout 500 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" >> $out/bench_default.err
tail -3 $out/pytest.txt; grep "==" $out/exp_c*.txt; tail -3 $out/bench_default.err
