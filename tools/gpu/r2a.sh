# round 2, call A: TMA-staged ROIAlign -- parity, microbench (staged vs gather), ncu capture; then the round-1 A/Bs that never ran.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "== parity (staged kernel)"; timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== parity (gather kernel)"; MRCNN_ROIALIGN=gather timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k roialign 2>&1 | tail -2
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,300,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
echo "== microbench staged"; timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/ra_staged.json 2>&1 | tail -6
echo "== microbench staged slot 32"; MRCNN_ROIALIGN_SLOT_PX=32 timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/ra_staged32.json 2>&1 | tail -6
echo "== microbench gather"; MRCNN_ROIALIGN=gather timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/ra_gather.json 2>&1 | tail -6
echo "== ncu staged b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2a_roialign_tma -f \
  python tools/bench_roialign.py --case "nhwc_f16,8,1000,7" --iters 3 --out gpurun_out/ra_ncu.json > gpurun_out/ncu_r2a.log 2>&1; tail -2 gpurun_out/ncu_r2a.log
echo "== pipeline bench (staged roialign)"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -2 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('stage_ms'))
print({k:v['ms_per_step'] for k,v in d['kernel_classes'].items()})
PY
echo "== stage split A/B"; timeout 900 python tools/ab_stage_split.py 2>&1 | tee gpurun_out/stage_split.txt | tail -8
echo "== BN sweep"; bash tools/gpu/bn_sweep.sh 2>&1 | tail -60
