#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2q; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tuple or golden_vectors or suffix_array or families" > $out/pytest_quick.txt 2>&1; echo "rc=$?" >> $out/pytest_quick.txt
BWTS_B200_TRACE=1 timeout 150 python tests/gpu_experiments.py C4 base 14:16 14:8 20:1 > $out/exp_c4.txt 2> $out/exp_c4_trace.txt
timeout 60 python tests/gpu_experiments.py C2 base 14:32 > $out/exp_c2.txt 2>&1
timeout 60 python tests/gpu_experiments.py C5 base > $out/exp_c5.txt 2>&1
timeout 100 python tests/gpu_experiments.py C6 base > $out/exp_c6.txt 2>&1
tail -3 $out/pytest_quick.txt; grep -A6 "==" $out/exp_c*.txt | grep -E "==|tuple|local_sort|rerank"
