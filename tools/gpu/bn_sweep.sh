# N-tile sweep on the layers that use BN = 256 (MRCNN_CONV_BN knob): per-layer table at 256 / 128 / 64
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -2
for bn in 0 128 64; do
  MRCNN_CONV_BN=$bn timeout 400 python tools/bench_conv_layers.py --only "${ONLY:-res4 2,res5 2,res3 2c,res4 sc,res5 sc,fpn lat}" --reps 9 2>&1 \
    | grep -v "^sum" | sed "s/^/bn=$bn /" | tee -a gpurun_out/bn_sweep.txt
done
