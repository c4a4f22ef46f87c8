# round 2, call M: two-stream streaming pipeline -- correctness (streaming == blocking, bit for bit) and throughput A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
  MRCNN_STREAM_HEADS=$v timeout 900 python bench.py --config A --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_heads$v.json 2> gpurun_out/bench_heads$v.err; tail -1 gpurun_out/bench_heads$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_heads$v.json')); print('heads=$v value', round(d['value'],1), 'blocking', round(d['blocking_value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), d['clocks'])"
done
