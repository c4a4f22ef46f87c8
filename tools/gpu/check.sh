# Run on a B200 box from the repo root (e.g. gpurun -- 'bash tools/gpu/check.sh'): GPU parity tests, smoke, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['stage_ms'], d['cpu_baseline']['value'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernel_classes'].items()})"
tail -3 gpurun_out/bench.err
