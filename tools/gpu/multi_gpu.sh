# N-GPU bench line (run with gpurun --gpus N -- 'bash tools/gpu/multi_gpu.sh N'): the driver's launch, 127.0.0.1 rendezvous
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print(d['n_gpus'], 'GPUs:', round(d['value'],1), 'streaming', round(d['streaming_device_value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), 'gather_verified', d['gather_verified'])"
