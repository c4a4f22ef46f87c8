mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1h.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"; tail -5 gpurun_out/bench_r1h.err
ncu --set full --warp-sampling-interval 1 --clock-control none --cache-control none --import-source on -k regex:conv_gemm_kernel -s 520 -c 1 -o gpurun_out/prof_conv1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h1.log 2>&1
ncu --set full --warp-sampling-interval 1 --clock-control none --cache-control none --import-source on -k regex:conv_gemm_kernel -s 627 -c 1 -o gpurun_out/prof_lat2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h2.log 2>&1
ncu --set full --warp-sampling-interval 1 --clock-control none --cache-control none --import-source on -k regex:det_finalize -s 4 -c 1 -o gpurun_out/prof_detfin -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h3.log 2>&1
ls -la gpurun_out/*.ncu-rep
