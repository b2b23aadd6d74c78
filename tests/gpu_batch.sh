#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2d; mkdir -p $out
timeout 500 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
BWTS_B200_TRACE=1 timeout 150 python tests/gpu_experiments.py C4 base 14:8 14:1 15:1 > $out/exp_c4.txt 2> $out/exp_c4_trace.txt
timeout 90 python tests/gpu_experiments.py C2 base 14:1 15:1 > $out/exp_c2.txt 2>&1
timeout 120 python tests/gpu_experiments.py C5 base 14:1 15:1 > $out/exp_c5.txt 2>&1
timeout 120 python tests/gpu_experiments.py C3 base > $out/exp_c3.txt 2>&1
tail -3 $out/pytest.txt; grep "==" $out/exp_c*.txt
