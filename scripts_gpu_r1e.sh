mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1e_2gpu.json 2> gpurun_out/bench_r1e_2gpu.err; tail -c 1500 gpurun_out/bench_r1e_2gpu.json; tail -5 gpurun_out/bench_r1e_2gpu.err
