mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/bench_conv_layers.py > gpurun_out/conv_layers_r1c.txt 2>&1; head -30 gpurun_out/conv_layers_r1c.txt; tail -2 gpurun_out/conv_layers_r1c.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 3500 gpurun_out/bench_r1c.json; tail -5 gpurun_out/bench_r1c.err
