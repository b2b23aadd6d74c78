#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
N=${1:-2}
out=gpurun_out/r3i_n$N; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $out/bench_c5_n$N.json 2> $out/bench_c5_n$N.err; echo "bench rc=$?" >> $out/bench_c5_n$N.err
tail -n 2 $out/bench_c5_n$N.err; grep -c '^{' $out/bench_c5_n$N.json
