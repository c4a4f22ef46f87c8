mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in 1 2; do echo "=== MRCNN_CONV_CTAS=$c"; MRCNN_CONV_CTAS=$c timeout 300 python tools/bench_conv_layers.py > gpurun_out/conv_layers_ctas$c.txt 2>&1; head -16 gpurun_out/conv_layers_ctas$c.txt; tail -1 gpurun_out/conv_layers_ctas$c.txt; done
for c in 1 2; do MRCNN_CONV_CTAS=$c timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ctas$c.json 2> gpurun_out/bench_ctas$c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ctas$c.json')); print('ctas $c', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'])"; tail -3 gpurun_out/bench_ctas$c.err; done
