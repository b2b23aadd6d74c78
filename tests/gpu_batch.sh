#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2p; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or families or random_small or binned_emit or lyndon or periodic or adversarial or generated_medium" > $out/pytest_quick.txt 2>&1; echo "rc=$?" >> $out/pytest_quick.txt
BWTS_B200_TRACE=1 timeout 100 python tests/gpu_experiments.py C4 base 9:3 > $out/exp_c4.txt 2> $out/exp_c4_trace.txt
timeout 60 python tests/gpu_experiments.py C2 base > $out/exp_c2.txt 2>&1
timeout 60 python tests/gpu_experiments.py C5 base > $out/exp_c5.txt 2>&1
timeout 100 python tests/gpu_experiments.py C3 base > $out/exp_c3.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -3 $out/pytest_quick.txt; tail -3 $out/pytest.txt; grep -A12 "==" $out/exp_c*.txt | grep -E "==|lyndon|emit"
