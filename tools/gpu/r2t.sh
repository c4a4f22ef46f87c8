mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 2 3; do
  MRCNN_ROIALIGN_CTAS=$v timeout 600 python bench.py --config A --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t$v.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_t$v.json')); print('ctas=$v', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['stage_ms'].items() if 'ROI' in k}, round(d['kernel_classes']['roialign']['ms_per_step'],4))"
done
