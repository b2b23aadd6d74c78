#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r3g; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not largest and not c6 and not c5" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 200 python tests/gpu_experiments.py C3 base 21:1 > $out/exp_c3.txt 2>&1
timeout 150 python tests/gpu_stress.py 60 16 > $out/stress.txt 2>&1
tail -n 3 $out/pytest.txt; grep "^==" $out/exp_*.txt; grep "init_keys\|radix_hist" $out/exp_c3.txt; tail -n 2 $out/stress.txt
