#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2k; mkdir -p $out
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 100 python tests/gpu_experiments.py C3 base 17:2 17:1 > $out/exp_c3.txt 2>&1
timeout 100 python tests/gpu_experiments.py C4 base 17:2 > $out/exp_c4.txt 2>&1
timeout 500 python bench.py > $out/bench_c4.json 2> $out/bench_c4.err; echo "bench rc=$?" >> $out/bench_c4.err
timeout 200 python bench.py --workload C2 --no-cli > $out/bench_c2.json 2> $out/bench_c2.err
timeout 300 python bench.py --workload C3 --no-cli --steps 3 > $out/bench_c3.json 2> $out/bench_c3.err
timeout 300 python bench.py --workload C3F --no-cli --no-cpu --steps 2 --warmup 1 > $out/bench_c3f.json 2> $out/bench_c3f.err
timeout 200 python bench.py --workload C1 --no-cli > $out/bench_c1.json 2> $out/bench_c1.err
python - <<'PY'
import sys
sys.path.insert(0, "tests")
import helpers
open("/dev/shm/c4.bin", "wb").write(helpers.Generator().make("dna", 4, 1 << 30))
PY
for i in 1 2; do
( time BWTS_B200_TIMINGS=1 bijective-bwt_b200/bin/mk_bwts /dev/shm/c4.bin /dev/shm/c4.bwts ) > $out/cli_fwd_$i.txt 2>&1
( time BWTS_B200_TIMINGS=1 bijective-bwt_b200/bin/unbwts /dev/shm/c4.bwts /dev/shm/c4.back ) > $out/cli_inv_$i.txt 2>&1
done
cmp /dev/shm/c4.bin /dev/shm/c4.back && echo "cli round trip ok" >> $out/cli_inv_2.txt
tail -3 $out/pytest.txt; grep "==" $out/exp_c*.txt; tail -2 $out/bench_c4.err
