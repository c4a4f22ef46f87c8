mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1m.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['stage_ms'])"; tail -3 gpurun_out/bench_r1m.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q > gpurun_out/sanitizer_layers.log 2>&1; tail -15 gpurun_out/sanitizer_layers.log
