mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in 0 1 0 1; do
  MRCNN_CONV_VGROUP=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vg$v.json 2> gpurun_out/bench_vg$v.err || tail -3 gpurun_out/bench_vg$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_vg$v.json')); print('vgroup $v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
