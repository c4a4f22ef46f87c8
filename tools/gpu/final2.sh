# end-of-round evidence after the fused expansion + reduction kernel: GPU suite, smoke, bench lines, launch list, memcheck
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
for c in A small single; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -1 gpurun_out/bench_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$c.json'))
print('$c', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],3), 'launches', d['gpu_launches'])
print(d.get('e2e_agreement'))
PY
done
timeout 200 python bench.py --config A --precise-masks 0 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_A_1term.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_A_1term.json')); print('A 1-term masks', d['value'], d['e2e']['value'])"
for c in small single; do
  MRCNN_CONV_FUSE=0 timeout 200 python bench.py --config $c --quick --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c unfused', round(d['value'],1), round(d['e2e']['value'],1))"
done
N=$(python - <<'PY'
import sys; sys.path.insert(0, '.')
import bench, maskrcnn_b200 as m, torch
wl = bench.PipelineWorkload(m, torch, 0, 8, 0, 1)
print(wl.launches_per_step())
PY
)
echo "launches per step: $N"
SKIP=$((N * 4))
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none -s $SKIP -c $N --csv --log-file gpurun_out/launches_r2k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_n0.log 2>&1; tail -1 gpurun_out/ncu_n0.log | cut -c1-200
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_pipeline_gpu.py tests/test_conv_gpu.py -m gpu -x -q > gpurun_out/memcheck_final2.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_final2.log
