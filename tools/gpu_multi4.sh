#!/bin/bash
# 4-GPU call: exchange variants at 4 ranks (loads/stores, TMA, multicast with the unroll / grid knobs), bench with multicast on
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 tools/peer_pieces.py 2> gpurun_out/pieces_4gpu.err | grep '^{' | tee gpurun_out/pieces_4gpu.json
for mc in 0 1; do
GSPLAT_B200_MULTICAST=$mc timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_4gpu_mc$mc.json 2> gpurun_out/bench_4gpu_mc$mc.err; echo "bench N=4 mc=$mc rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_4gpu_mc$mc.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'allreduce',d.get('allreduce',{}).get('ms'), (d.get('exchange_check') or {}).get('bitwise_same_on_all_ranks'), (d.get('exchange_check') or {}).get('max_rel_err_vs_nccl'))
PY
done
