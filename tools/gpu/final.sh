# end-of-round evidence: GPU suite, smoke, bench lines (A / small / single / 1-term masks), ROIAlign evidence, memcheck of the pipeline tests
mkdir -p gpurun_out
bash tools/gpu/bench_configs.sh
bash tools/gpu/roialign_evidence.sh
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_pipeline_gpu.py tests/test_layers_gpu.py -m gpu -x -q > gpurun_out/memcheck_final.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_final.log
