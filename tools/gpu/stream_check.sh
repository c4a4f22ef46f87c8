# streaming API tests + bench with both tail-wait variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tailread.json 2> gpurun_out/bench_tailread.err
MRCNN_CONV_TAIL_FULL_WAIT=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tailfull.json 2> gpurun_out/bench_tailfull.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tailread2.json 2>> gpurun_out/bench_tailread.err
python - <<'PY'
import json
for n in ("tailread", "tailfull", "tailread2"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "sync", round(d["e2e"]["sync_value"], 1), "conv", round(d["roofline"]["achieved"], 1), d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/bench_tailread.err
