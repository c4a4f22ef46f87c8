# round-end evidence: full GPU suite, smoke, bench (with CPU baseline), warm ncu launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_roialign']['frac'], d['cpu_baseline']['value'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none -s 628 -c 157 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
wc -l gpurun_out/launches_final.csv
