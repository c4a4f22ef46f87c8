# quick check after a ROIAlign change: parity of every path + the four headline microbench cases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/bench_roialign.py --case "nhwc_f16,1,1000,7;nhwc_f16,8,1000,7;nhwc_f16,8,1000,14;nhwc_f16,64,1000,7" --out gpurun_out/quick.json 2>&1 | tail -4
