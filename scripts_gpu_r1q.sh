mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_conv_layers.py > gpurun_out/conv_layers_r1q.txt 2>&1; head -30 gpurun_out/conv_layers_r1q.txt; tail -1 gpurun_out/conv_layers_r1q.txt
MRCNN_CONV_CTAS=2 timeout 300 python tools/bench_conv_layers.py --only "res2,res3,fpn lat" > gpurun_out/conv_layers_r1q_c2.txt 2>&1; cat gpurun_out/conv_layers_r1q_c2.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1q.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms']['Backbone+FPN+RPN'])"; tail -2 gpurun_out/bench_r1q.err
