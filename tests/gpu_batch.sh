#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2l; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_vectors or families or random_small or tuple or cta_local or local_sort or binned or context_reuse or two_contexts or adversarial" > $out/pytest_quick.txt 2>&1; echo "rc=$?" >> $out/pytest_quick.txt
timeout 100 python tests/gpu_experiments.py C4 base > $out/exp_c4.txt 2>&1
timeout 100 python tests/gpu_experiments.py C3 base > $out/exp_c3.txt 2>&1
timeout 60 python tests/gpu_experiments.py C2 base > $out/exp_c2.txt 2>&1
timeout 60 python tests/gpu_experiments.py C5 base > $out/exp_c5.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -3 $out/pytest_quick.txt; tail -3 $out/pytest.txt; grep -A3 "==" $out/exp_c*.txt | grep -E "==|rerank"
