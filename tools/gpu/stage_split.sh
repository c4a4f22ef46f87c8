# half-batch stage split experiment: regression first, then the in-process A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python tools/ab_stage_split.py 2>&1 | tee gpurun_out/stage_split.txt | tail -8
