mkdir -p gpurun_out
for k in sort_decode nms_mask proposal_resolve det_finalize; do
ncu --set full --warp-sampling-interval 1 --clock-control none --cache-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_n_$k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_n_$k.log 2>&1
done
ls gpurun_out/prof_n_*
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1n.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['stage_ms'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"; tail -3 gpurun_out/bench_r1n.err
