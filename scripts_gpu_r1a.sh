set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -c 3000 gpurun_out/bench_r1a.json; tail -5 gpurun_out/bench_r1a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:roialign_chw -s 2 -c 2 -o gpurun_out/prof_roialign_r1a -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
