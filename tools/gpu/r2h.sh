# round 2, call H: CHW staged path parity + microbench, new NHWC tests, bench --workload roialign
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -4
CASES="chw_f32,1,1000,7;chw_f32,8,1000,7;chw_f32,8,1000,14;nhwc_f16,8,1000,7"
echo "== microbench staged"; timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rh.json 2>&1 | tail -4
echo "== microbench gather"; MRCNN_ROIALIGN=gather timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rh_g.json 2>&1 | tail -4
echo "== ncu chw b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2h_roialign_chw -f \
  python tools/bench_roialign.py --case "chw_f32,8,1000,7" --iters 3 --out gpurun_out/rh_ncu.json > gpurun_out/ncu_r2h.log 2>&1; tail -2 gpurun_out/ncu_r2h.log
