# warm launch list of one pipeline step (all kernels) + full ncu captures of the proposal-stage kernels
mkdir -p gpurun_out
# how many launches precede the 4th step: count them from the library (warm-up 3 + 1 launch-count step)
N=$(python - <<'PY'
import sys; sys.path.insert(0, '.')
import bench, maskrcnn_b200 as m, torch
wl = bench.PipelineWorkload(m, torch, 0, 8, 0, 1)
print(wl.launches_per_step())
PY
)
echo "launches per step: $N"
SKIP=$((N * 4))
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none -s $SKIP -c $N --csv --log-file gpurun_out/launches_r2n.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_n0.log 2>&1; tail -1 gpurun_out/ncu_n0.log
for k in nms_mask_kernel sel_hist_kernel sort_decode_kernel proposal_resolve_kernel; do
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s 8 -c 1 -o gpurun_out/r2n_$k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_n1.log 2>&1; tail -1 gpurun_out/ncu_n1.log
done
