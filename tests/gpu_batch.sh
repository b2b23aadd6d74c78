#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2o; mkdir -p $out
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 60 python tests/gpu_experiments.py C2 base > $out/exp_c2.txt 2>&1
timeout 60 python tests/gpu_experiments.py C5 base > $out/exp_c5.txt 2>&1
timeout 100 python tests/gpu_experiments.py C4 base > $out/exp_c4.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 3 --warmup 2 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench rc=$?" >> $out/bench_n2.err
timeout 100 python bench.py --workload C1 --no-cli > $out/bench_c1.json 2> $out/bench_c1.err
tail -3 $out/pytest.txt; grep "==" $out/exp_c*.txt; tail -3 $out/bench_n2.err; cut -c1-200 $out/bench_n2.json
