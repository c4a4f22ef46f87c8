mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1g.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"; tail -5 gpurun_out/bench_r1g.err
