# evidence for DESIGN.md section 8 (chained ResNet stages): tile-count probe, per-role stall statistics, interleaved A/B
mkdir -p gpurun_out
bash tools/gpu/quant_probe.sh > gpurun_out/quant_probe.txt 2>&1
MRCNN_CHAIN=1 MRCNN_CHAIN_STAGES=4 timeout 120 python tools/chain_stats.py > gpurun_out/chain_stats_res4.txt 2>&1
VARIANTS="0:0 1:4:0 1:4:2 1:14:0 1:15:0" timeout 300 python tools/ab_chain.py > gpurun_out/ab_chain.txt 2>&1
NCTX=2 timeout 200 python tools/two_ctx.py > gpurun_out/two_ctx.txt 2>&1
for f in quant_probe chain_stats_res4 ab_chain two_ctx; do echo "== $f"; tail -n 6 gpurun_out/$f.txt; done
