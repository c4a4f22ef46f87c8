mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q -k "chained" 2>&1 | tail -3
for len in ${LENS:-1 96}; do
  MRCNN_CHAIN=1 MRCNN_CHAIN_STAGES=${MASK:-4} MRCNN_CHAIN_MAXLEN=$len timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_len$len.json 2> gpurun_out/bench_len$len.err || tail -5 gpurun_out/bench_len$len.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_len$len.json"))
print("maxlen $len", round(d["value"], 1), "backbone ms", round(d["stage_ms"]["Backbone+FPN+RPN"], 3), "conv TF/s", round(d["roofline"]["achieved"], 1), d["clocks"]["sm_mhz"])
PY
done
