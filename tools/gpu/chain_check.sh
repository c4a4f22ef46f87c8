mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q -k "chained" 2>&1 | tail -5
MASKS="${MASKS:-0 2 4 8 14 15}" bash tools/gpu/chain_stages.sh
MRCNN_CHAIN=1 timeout 300 python tools/chain_stats.py 2>&1 | grep -v "^   epi:res\|^   epi:bulk"
