#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r3j; mkdir -p $out
timeout 100 python tests/gpu_experiments.py C5 base 23:96 23:128 23:48 23:86 > $out/exp_c5.txt 2>&1
grep "^==\|fwd emit" $out/exp_c5.txt
