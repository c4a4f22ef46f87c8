mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
CASES="nhwc_f16,1,1000,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7"
for v in "CTAS=2" "CTAS=3"; do
echo "== microbench $v"; env MRCNN_ROIALIGN_$v timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/rr.json 2>&1 | tail -4
done
