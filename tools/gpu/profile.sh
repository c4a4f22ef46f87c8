# ncu evidence of one pipeline step: warm launch list + full captures (CTA-pair conv kernel on the RPN 3x3 conv, ROIAlign).
mkdir -p gpurun_out
# launch list of one step (warm caches) + one full capture each of the CTA-pair conv kernel (RPN shared conv on P2), the 1-CTA staged-epilogue conv (res4 2c) and ROIAlign
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none -s 628 -c 157 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l0.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:conv_gemm_kernel -s 633 -c 1 -o gpurun_out/prof_r1l_conv_rpn_p2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:roialign_nhwc -s 8 -c 1 -o gpurun_out/prof_r1l_roialign -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
