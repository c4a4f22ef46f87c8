mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
for h in 1 0; do MRCNN_NO_REVERSE=$h timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rev$h.json 2> gpurun_out/bench_rev$h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_rev$h.json')); print('NO_REVERSE=$h', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms']['Backbone+FPN+RPN'])"; tail -2 gpurun_out/bench_rev$h.err; done
