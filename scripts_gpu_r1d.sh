mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/bench_conv_layers.py > gpurun_out/conv_layers_r1d.txt 2>&1; head -24 gpurun_out/conv_layers_r1d.txt; tail -1 gpurun_out/conv_layers_r1d.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1d.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"; tail -5 gpurun_out/bench_r1d.err
