#!/bin/bash
# One gpurun call's worth of measurements (scratch output under gpurun_out/); edited per call.
# Every command runs under its own timeout: a hung kernel must not eat the box's time limit.
out=gpurun_out/r2g; mkdir -p $out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lyndon or golden_vectors or periodic" > $out/pytest_quick.txt 2>&1; echo "rc=$?" >> $out/pytest_quick.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for wl in C4 C3 C2; do
  P="python tests/gpu_profile_target.py $wl"
  timeout 120 $P > $out/plain_$wl.log 2>&1 && timeout 400 ncu --metrics $M --clock-control none --csv --log-file $out/launches_$wl.csv $P > $out/ncu_list_$wl.log 2>&1
done
P="python tests/gpu_profile_target.py C4"
for pat in k_rerank k_tuple_round k_tuple_apply k_local_sort_warp k_inv_walk_stage k_inv_place_copy k_build_keys k_duval_chunks k_scatter_pairs k_scatter_bytes; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -c 2 -o $out/full_C4_$pat $P > $out/ncu_full_C4_$pat.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_onesweep_pass -s 4 -c 2 -o $out/full_C4_k_onesweep_pass $P > $out/ncu_full_C4_k_onesweep_pass.log 2>&1
P="python tests/gpu_profile_target.py C3"
for pat in k_local_sort_cta_radix k_sufmin_reduce_cta; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 2 -o $out/full_C3_$pat $P > $out/ncu_full_C3_$pat.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rerank -s 12 -c 2 -o $out/full_C3_k_rerank $P > $out/ncu_full_C3_k_rerank.log 2>&1
for f in $out/full_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $f --page details --csv > ${b}_details.csv 2>/dev/null
done
ls -la $out | head -70
# keep the source-level reports of the two top kernels only (gpurun_out is capped at 64 MiB)
for f in $out/full_*.ncu-rep; do case $f in *C4_k_rerank*|*k_tuple_round*) ;; *) rm -f $f;; esac; done
# CLI breakdown on the 1 GiB DNA file
python - <<'PY'
import sys
sys.path.insert(0, "tests")
import helpers
open("/dev/shm/c4.bin", "wb").write(helpers.Generator().make("dna", 4, 1 << 30))
PY
( time BWTS_B200_TIMINGS=1 bijective-bwt_b200/bin/mk_bwts /dev/shm/c4.bin /dev/shm/c4.bwts ) > $out/cli_fwd.txt 2>&1
( time BWTS_B200_TIMINGS=1 bijective-bwt_b200/bin/unbwts /dev/shm/c4.bwts /dev/shm/c4.back ) > $out/cli_inv.txt 2>&1
( time bijective-bwt_b200/bin/mk_bwts /dev/shm/c4.bin /dev/shm/c4.bwts2 ) > $out/cli_fwd_plain.txt 2>&1
cmp /dev/shm/c4.bin /dev/shm/c4.back && echo "cli round trip ok" >> $out/cli_inv.txt
du -sh $out; tail -2 $out/pytest_quick.txt
