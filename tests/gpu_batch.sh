#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2h; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lyndon or cta_local or golden_vectors or families or adversarial or periodic or suffix_array" > $out/pytest_quick.txt 2>&1; echo "rc=$?" >> $out/pytest_quick.txt
BWTS_B200_TRACE=1 timeout 120 python tests/gpu_experiments.py C3 base 18:1 17:1 > $out/exp_c3.txt 2> $out/exp_c3_trace.txt
timeout 100 python tests/gpu_experiments.py C4 base 17:1 > $out/exp_c4.txt 2>&1
timeout 100 python tests/gpu_experiments.py C3F base 18:1 > $out/exp_c3f.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -3 $out/pytest_quick.txt; tail -3 $out/pytest.txt; grep "==" $out/exp_c*.txt
