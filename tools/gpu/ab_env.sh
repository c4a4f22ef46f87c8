# A/B of one environment knob on the batch-8 bench (device-resident value, conv class ms): bash tools/gpu/ab_env.sh VAR A B [reps]
VAR=$1; A=$2; B=$3; REPS=${4:-2}
mkdir -p gpurun_out
for r in $(seq 1 $REPS); do
  for v in $A $B; do
    env $VAR=$v timeout ${AB_TIMEOUT:-300} python bench.py --quick --steps 20 --warmup 3 > gpurun_out/ab_${VAR}_${v}_$r.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/ab_${VAR}_${v}_$r.json')); print('$VAR=$v', 'run $r', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', 'conv', round(d['kernel_classes']['conv_gemm_tcgen05']['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'])"
  done
done
