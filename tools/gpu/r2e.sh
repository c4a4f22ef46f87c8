# round 2, call E: staged ROIAlign v6 (planner + issuer warps, chunk-granular ring, optional L2 prefetch): parity + variants + ncu
mkdir -p gpurun_out
echo "== parity (staged kernel)"; timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== parity (staged, 3 CTAs, ahead 2)"; MRCNN_ROIALIGN_CTAS=3 MRCNN_ROIALIGN_AHEAD=2 timeout 600 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k roialign 2>&1 | tail -2
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
for v in "CTAS=2" "CTAS=3" "CTAS=2 MRCNN_ROIALIGN_AHEAD=2" "CTAS=2 MRCNN_ROIALIGN_AHEAD=1" "CTAS=3 MRCNN_ROIALIGN_AHEAD=1"; do
  echo "== microbench staged $v"; env MRCNN_ROIALIGN_$v timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/re.json 2>&1 | tail -4
done
echo "== ncu staged b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2e_roialign_tma -f \
  python tools/bench_roialign.py --case "nhwc_f16,8,1000,7" --iters 3 --out gpurun_out/re_ncu.json > gpurun_out/ncu_r2e.log 2>&1; tail -2 gpurun_out/ncu_r2e.log
