#!/bin/bash
# 8-GPU call (gpurun --gpus 8): pieces of the exchange, bench at 8 and 4
set -u
TAG=${1:-m8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/peer_pieces.py 2> gpurun_out/pieces_8gpu_$TAG.err | grep '^{' | tee gpurun_out/pieces_8gpu_$TAG.json
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_$TAG.json'))
    print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'wall',round(d['wall_frames_per_s'],1),'allreduce',d.get('allreduce',{}).get('ms'))
    print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
    print({k:v for k,v in (d.get('exchange_check') or {}).items() if k != 'what'})
except Exception as e: print('no bench line', e)
PY
done
