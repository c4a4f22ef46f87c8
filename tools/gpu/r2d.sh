# round 2, call D: staged ROIAlign v4 (producer-written descriptors, lean consumer loop): parity + variants + ncu
mkdir -p gpurun_out
echo "== parity (staged kernel)"; timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== parity (staged, 5 slots)"; MRCNN_ROIALIGN_SLOTS=5 timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k roialign 2>&1 | tail -2
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
for v in "SLOTS=8" "SLOTS=5" "SLOTS=8 MRCNN_ROIALIGN_SLOT_PX=16" "SLOTS=5 MRCNN_ROIALIGN_SLOT_PX=16"; do
  echo "== microbench staged $v"; env MRCNN_ROIALIGN_$v timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rd.json 2>&1 | tail -4
done
echo "== ncu staged b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2d_roialign_tma -f \
  python tools/bench_roialign.py --case "nhwc_f16,8,1000,7" --iters 3 --out gpurun_out/rd_ncu.json > gpurun_out/ncu_r2d.log 2>&1; tail -2 gpurun_out/ncu_r2d.log
