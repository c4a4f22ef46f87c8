mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_conv_layers.py > gpurun_out/conv_layers_r1k.txt 2>&1; tail -1 gpurun_out/conv_layers_r1k.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1k.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['stage_ms'], d['clocks'])"; tail -3 gpurun_out/bench_r1k.err
