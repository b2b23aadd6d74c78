#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
#   gpurun --timeout 3300 -- 'bash tests/gpu_batch.sh'
out=gpurun_out/batch; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
timeout 500 python bench.py > $out/bench_c4.json 2> $out/bench_c4.err; echo "bench rc=$?" >> $out/bench_c4.err
for wl in C2 C3 C3F C6 C1; do
  timeout 300 python bench.py --workload $wl > $out/bench_$wl.json 2> $out/bench_$wl.err; echo "bench $wl rc=$?" >> $out/bench_$wl.err
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for wl in C4 C3 C2; do
  P="python tests/gpu_profile_target.py $wl"
  timeout 120 $P > $out/plain_$wl.log 2>&1 && timeout 400 ncu --metrics $M --clock-control none --csv --log-file $out/launches_$wl.csv $P > $out/ncu_list_$wl.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?" >> $out/smoke.txt
tail -n 3 $out/pytest.txt; tail -n 1 $out/bench_*.err; tail -n 2 $out/smoke.txt
# afterwards, here: cp the launch lists to profiles/r02_launches_<wl>.csv, python profiles/summarize_ncu_launches.py,
# python profiles/make_traffic_json.py
