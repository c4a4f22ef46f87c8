# quick regression + bench (no CPU baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'])"; tail -2 gpurun_out/bench_quick.err
