# round 2, call C: staged ROIAlign v3 (rolled loops, L2 prefetch pass): parity + variants + ncu
mkdir -p gpurun_out
echo "== parity (staged kernel)"; timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
for v in "AHEAD=0" "AHEAD=2" "AHEAD=4" "AHEAD=8" "AHEAD=2 MRCNN_ROIALIGN_SLOTS=5" "AHEAD=4 MRCNN_ROIALIGN_SLOTS=5"; do
  echo "== microbench staged $v"; env MRCNN_ROIALIGN_$v timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rc.json 2>&1 | tail -4
done
echo "== ncu staged b8 R1000 P7 (ahead 2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2c_roialign_tma -f \
  python tools/bench_roialign.py --case "nhwc_f16,8,1000,7" --iters 3 --out gpurun_out/rc_ncu.json > gpurun_out/ncu_r2c.log 2>&1; tail -2 gpurun_out/ncu_r2c.log
