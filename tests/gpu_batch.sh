#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2j; mkdir -p $out
nvidia-smi -L > $out/box.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dealer or blocks_equal or lyndon_scan or timings" > $out/pytest_multi.txt 2>&1; echo "rc=$?" >> $out/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 2 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench rc=$?" >> $out/bench_n2.err
timeout 300 python bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $out/bench_ref_n2.json 2> $out/bench_ref_n2.err
tail -3 $out/pytest_multi.txt; tail -5 $out/bench_n2.err; cut -c1-600 $out/bench_n2.json
