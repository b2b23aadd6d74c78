#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2n; mkdir -p $out
nvidia-smi -L > $out/box.txt; nvidia-smi topo -m >> $out/box.txt 2>&1
N=$(nvidia-smi -L | wc -l)
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dealer or blocks_equal" > $out/pytest_multi.txt 2>&1; echo "rc=$?" >> $out/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 3 --warmup 2 > $out/bench_n$N.json 2> $out/bench_n$N.err; echo "bench rc=$?" >> $out/bench_n$N.err
tail -3 $out/pytest_multi.txt; tail -4 $out/bench_n$N.err; cut -c1-300 $out/bench_n$N.json
