set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python __graft_entry__.py smoke 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -c 4000 gpurun_out/bench_r1b.json; tail -20 gpurun_out/bench_r1b.err
