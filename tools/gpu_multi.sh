#!/bin/bash
# multi-GPU call (gpurun --gpus N): peer exchange tests vs NCCL, pieces of the exchange, bench at N
set -u
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_peer_exchange.py -q --timeout 300 > gpurun_out/pytest_peer_${N}gpu_$TAG.log 2>&1; echo "peer pytest rc=$?"
tail -6 gpurun_out/pytest_peer_${N}gpu_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/peer_pieces.py 2> gpurun_out/pieces_${N}gpu_$TAG.err | grep '^{' | tee gpurun_out/pieces_${N}gpu_$TAG.json
for tma in 0 1; do
  GSPLAT_B200_PEER_TMA=$tma timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_tma${tma}_$TAG.json 2> gpurun_out/bench_${N}gpu_tma${tma}_$TAG.err; echo "bench N=$N tma=$tma rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_tma${tma}_$TAG.json'))
    print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'allreduce',d.get('allreduce',{}).get('ms'),'check',d.get('exchange_check'))
except Exception as e: print('no bench line', e)
PY
done
nvidia-smi topo -m > gpurun_out/topo_${N}gpu_$TAG.txt 2>&1
