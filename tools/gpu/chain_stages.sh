# backbone time with exactly one ResNet stage chained (bit s of MRCNN_CHAIN_STAGES = res(s+2))
mkdir -p gpurun_out
for m in ${MASKS:-0 1 2 4 8 15}; do
  MRCNN_CHAIN=1 MRCNN_CHAIN_STAGES=$m timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline > gpurun_out/bench_mask$m.json 2> gpurun_out/bench_mask$m.err || tail -5 gpurun_out/bench_mask$m.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_mask$m.json"))
print("mask $m", round(d["value"], 1), "backbone ms", round(d["stage_ms"]["Backbone+FPN+RPN"], 3), "conv TF/s", round(d["roofline"]["achieved"], 1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
