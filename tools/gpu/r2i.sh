mkdir -p gpurun_out
CASES="chw_f32,1,1000,7;chw_f32,8,1000,7;chw_f32,8,1000,14"
echo "== microbench staged"; timeout 300 python tools/bench_roialign.py --case "$CASES" --out gpurun_out/ri.json 2>&1 | tail -4
echo "== ncu chw b8 R1000 P7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roialign_staged -s 3 -c 1 -o gpurun_out/r2i_roialign_chw -f \
  python tools/bench_roialign.py --case "chw_f32,8,1000,7" --iters 3 --out gpurun_out/ri_ncu.json > gpurun_out/ncu_r2i.log 2>&1; tail -2 gpurun_out/ncu_r2i.log
