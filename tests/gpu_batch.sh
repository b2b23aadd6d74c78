#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2x; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not largest and not c5 and not c6" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 200 python tests/gpu_experiments.py C4 base 21:1 1:26 > $out/exp_c4.txt 2>&1
timeout 100 python tests/gpu_experiments.py C2 base 21:1 > $out/exp_c2.txt 2>&1
timeout 100 python tests/gpu_experiments.py C5 base 21:1 1:25 > $out/exp_c5.txt 2>&1
timeout 100 python tests/gpu_experiments.py C6 base 1:26 > $out/exp_c6.txt 2>&1
timeout 150 python tests/gpu_stress.py 100 12 > $out/stress.txt 2>&1
tail -3 $out/pytest.txt; grep "^==" $out/exp_*.txt; grep "radix_hist\|init_keys\|inv_jump\|inv_walk\|inv_place" $out/exp_c4.txt $out/exp_c5.txt; tail -2 $out/stress.txt
