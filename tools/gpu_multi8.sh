#!/bin/bash
# 8-GPU call (gpurun --gpus 8): peer exchange tests at worlds 3 and 8, pieces of the exchange, bench at 8 (and 4)
set -u
TAG=${1:-m8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_peer_exchange.py -q --timeout 200 -k "3- or 8-" > gpurun_out/pytest_peer_8gpu_$TAG.log 2>&1; echo "peer pytest rc=$?"
tail -4 gpurun_out/pytest_peer_8gpu_$TAG.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/peer_pieces.py 2> gpurun_out/pieces_8gpu_$TAG.err | grep '^{' | tee gpurun_out/pieces_8gpu_$TAG.json
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_$TAG.json'))
    print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'wall',round(d['wall_frames_per_s'],1),'allreduce',d.get('allreduce',{}).get('ms'))
    print(d.get('exchange_check'))
except Exception as e: print('no bench line', e)
PY
done
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
