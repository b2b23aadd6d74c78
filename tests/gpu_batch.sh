#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
out=gpurun_out/r2r; mkdir -p $out
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 200 python tests/gpu_stress.py 150 11 > $out/stress.txt 2>&1; echo "stress rc=$?" >> $out/stress.txt
timeout 500 python bench.py > $out/bench_c4.json 2> $out/bench_c4.err; echo "bench rc=$?" >> $out/bench_c4.err
timeout 200 python bench.py --workload C2 --no-cli > $out/bench_c2.json 2> $out/bench_c2.err
timeout 300 python bench.py --workload C3 --no-cli --steps 3 > $out/bench_c3.json 2> $out/bench_c3.err
timeout 300 python bench.py --workload C6 --no-cli --no-cpu --steps 2 --warmup 1 > $out/bench_c6.json 2> $out/bench_c6.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_ref_c4.json 2> $out/bench_ref_c4.err
tail -3 $out/pytest.txt; tail -3 $out/stress.txt; tail -2 $out/bench_c4.err; tail -2 $out/bench_c6.err
