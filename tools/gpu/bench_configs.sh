# GPU suite + bench lines for configs A / small / single and the 1-term mask mode
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for c in A small single; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$c.json'))
print('$c', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],3))
print(d.get('e2e_agreement')); print(d['stage_ms'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernel_classes'].items()})
PY
done
timeout 600 python bench.py --config A --precise-masks 0 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_A_1term.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_A_1term.json')); print('A 1-term masks', d['value'], d['e2e']['value'])"
