mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --config A --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_p.json')); print(round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_step'],4) for k,v in d['kernel_classes'].items()})"
