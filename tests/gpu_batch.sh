#!/bin/bash
out=gpurun_out/r2t; mkdir -p $out
timeout 60 python tests/gpu_experiments.py C2 base "11:1!" "15:2" > $out/exp_c2.txt 2>&1
timeout 60 python tests/gpu_experiments.py C5 base "11:1!" "15:2" > $out/exp_c5.txt 2>&1
grep -A4 "==" $out/exp_c2.txt $out/exp_c5.txt | grep -E "==|inv_walk"
