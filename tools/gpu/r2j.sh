mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== parity SORT=1 (nhwc too)"; MRCNN_ROIALIGN_SORT=1 timeout 900 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k roialign 2>&1 | tail -2
CASES="chw_f32,1,1000,7;chw_f32,8,1000,7;chw_f32,8,1000,14;nhwc_f16,8,1000,7;nhwc_f16,1,1000,7;nhwc_f16,64,1000,7"
for v in "SORT=1" "SORT=0"; do
echo "== microbench staged $v"; env MRCNN_ROIALIGN_$v timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rj.json 2>&1 | tail -6
done
echo "== ncu chw b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2j_roialign_chw -f \
  python tools/bench_roialign.py --case "chw_f32,8,1000,7" --iters 3 --out gpurun_out/rj_ncu.json > gpurun_out/ncu_r2j.log 2>&1; tail -2 gpurun_out/ncu_r2j.log
