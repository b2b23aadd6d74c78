# experiment helper (not a test): sweep a tune key over values on a workload
# usage: bash tests/gpu_sweep.sh <key> "<values>" [workload] 
KEY=$1; VALS=$2; WL=${3:-C2}
for t in $VALS; do
  timeout 400 python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu --tune $KEY:$t > gpurun_out/sw_${WL}_${KEY}_$t.json 2>gpurun_out/sw_${WL}_${KEY}_$t.err || tail -3 gpurun_out/sw_${WL}_${KEY}_$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/sw_${WL}_${KEY}_$t.json'))
kc=d['kernel_classes']
print('$WL tune $KEY:$t', 'fwd_ms %.2f inv_ms %.2f' % (d['forward_ms'], d['inverse_ms']), 'onesweep GB/s %.0f frac %.3f' % (d['roofline']['achieved'], d['roofline']['frac']), ' '.join('%s=%.2f' % (k, v['ms_per_step']) for k, v in list(kc.items())[:7]))
PY
done
