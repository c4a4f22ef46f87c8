# round 2, call F: staged ROIAlign v8 (row-centric consumers: x-lerp once per row, immediate release)
mkdir -p gpurun_out
echo "== parity (default)"; timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in "CTAS=2 MRCNN_ROIALIGN_CW=4" "CTAS=4 MRCNN_ROIALIGN_CW=4" "CTAS=3 MRCNN_ROIALIGN_CW=4"; do
echo "== parity $v"; env MRCNN_ROIALIGN_$v timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k roialign 2>&1 | tail -1
done
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
for v in "CTAS=2 MRCNN_ROIALIGN_CW=7" "CTAS=2 MRCNN_ROIALIGN_CW=4" "CTAS=3 MRCNN_ROIALIGN_CW=4" "CTAS=4 MRCNN_ROIALIGN_CW=4"; do
  echo "== microbench staged $v"; env MRCNN_ROIALIGN_$v timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rf.json 2>&1 | tail -4
done
echo "== ncu staged b8 R1000 P7 default"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2f_roialign_tma -f \
  python tools/bench_roialign.py --case "nhwc_f16,8,1000,7" --iters 3 --out gpurun_out/rf_ncu.json > gpurun_out/ncu_r2f.log 2>&1; tail -2 gpurun_out/ncu_r2f.log
