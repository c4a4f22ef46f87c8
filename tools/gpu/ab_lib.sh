# A/B of two builds of the library on one box: lib_base.so (copy of the previous build, repo root) vs the in-tree build
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in base new; do
  if [ $v = base ]; then export MRCNN_LIB_PATH=$PWD/lib_base.so; else unset MRCNN_LIB_PATH; fi
  timeout 300 python tools/bench_conv_layers.py --only "${ONLY:-res4 2,res3 2c,res2 2c,res5 2c,res2 sc,fpn lat2}" --reps 9 2>&1 | grep -v "^sum\|^layer" | sed "s/^/$v /"
done
for v in base new base new; do
  if [ $v = base ]; then export MRCNN_LIB_PATH=$PWD/lib_base.so; else unset MRCNN_LIB_PATH; fi
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', round(d['value'],1), 'conv', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
