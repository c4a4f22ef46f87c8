# round 2, call B: staged ROIAlign v2 (templated P, x-lerp row reuse): parity + microbench variants + ncu
mkdir -p gpurun_out
echo "== parity (staged kernel)"; timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== parity (staged, 5 slots)"; MRCNN_ROIALIGN_SLOTS=5 timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k roialign 2>&1 | tail -2
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,300,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
echo "== microbench staged 8 slots x 2 CTAs"; timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rb_staged.json 2>&1 | tail -6
echo "== microbench staged 5 slots x 3 CTAs"; MRCNN_ROIALIGN_SLOTS=5 timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rb_staged5.json 2>&1 | tail -6
echo "== microbench staged 16 px slots"; MRCNN_ROIALIGN_SLOT_PX=16 timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rb_staged16.json 2>&1 | tail -6
echo "== ncu staged b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2b_roialign_tma -f \
  python tools/bench_roialign.py --case "nhwc_f16,8,1000,7" --iters 3 --out gpurun_out/rb_ncu.json > gpurun_out/ncu_r2b.log 2>&1; tail -2 gpurun_out/ncu_r2b.log
