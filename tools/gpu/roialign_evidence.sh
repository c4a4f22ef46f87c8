# ROIAlign evidence (configs[4]): DRAM bytes of every sweep case (ncu metrics pass), the full sweep through bench.py, one full ncu capture each of the pool-7 and pool-14 kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:roialign_staged --csv --log-file gpurun_out/ra_dram.csv \
  python tools/bench_roialign.py --iters 1 --out gpurun_out/ra_dram_sweep.json > gpurun_out/ra_dram.log 2>&1; tail -2 gpurun_out/ra_dram.log
python tools/roialign_dram.py gpurun_out/ra_dram.csv gpurun_out/ra_dram_sweep.json profiles/r2_roialign_dram.json && cp profiles/r2_roialign_dram.json gpurun_out/
timeout 900 python bench.py --workload roialign --steps 6 > gpurun_out/bench_roialign.json 2> gpurun_out/bench_roialign.err; tail -32 gpurun_out/bench_roialign.err
for c in "nhwc_f16,8,1000,7" "nhwc_f16,8,1000,14"; do
  n=$(echo $c | tr ',' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2k_$n -f python tools/bench_roialign.py --case "$c" --iters 3 --out gpurun_out/rk.json > gpurun_out/ncu_r2k.log 2>&1; tail -1 gpurun_out/ncu_r2k.log
done
