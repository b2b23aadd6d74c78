#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2v; mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned or cta or largest_length" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 200 python tests/gpu_experiments.py C3 base 7:3 6:56 > $out/exp_c3.txt 2>&1
timeout 200 python tests/gpu_experiments.py C4 base 7:3 6:40 6:48 6:56 > $out/exp_c4.txt 2>&1
timeout 100 python tests/gpu_experiments.py C2 base 7:3 6:48 6:56 > $out/exp_c2.txt 2>&1
timeout 200 python tests/gpu_experiments.py C3F base 7:3 > $out/exp_c3f.txt 2>&1
tail -3 $out/pytest.txt; grep "^==" $out/exp_*.txt
