#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/box.txt; nproc >> $out/box.txt; free -g | head -2 >> $out/box.txt
BWTS_B200_TRACE=1 python tests/gpu_experiments.py C4 base 12:1 10:32 10:128 10:64 "10:64,11:1!" "10:64,12:1,11:1!" > $out/exp_c4.txt 2> $out/exp_c4_trace.txt
python tests/gpu_experiments.py C2 base 12:1 10:32 "10:32,12:1" 10:64 > $out/exp_c2.txt 2>&1
python tests/gpu_experiments.py C3 base 10:32 10:64 > $out/exp_c3.txt 2>&1
python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1
python bench.py > $out/bench_default.json 2> $out/bench_default.err || python bench.py --tune 12:1 > $out/bench_default_oldinv.json 2> $out/bench_default_oldinv.err
tail -3 $out/pytest.txt; tail -3 $out/bench_default.err
