# full GPU suite, smoke, compute-sanitizer memcheck + racecheck of the custom-layer kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k "roialign or proposal" > gpurun_out/memcheck_r2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_r2.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_layers_gpu.py -m gpu -x -q -k "test_pyramid_roialign_nhwc_f16 or test_roialign_nhwc_f16_many" > gpurun_out/racecheck_r2.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_r2.log
