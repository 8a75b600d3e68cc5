#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu_r2final.json 2> gpurun_out/bench_2gpu_r2final.err; echo "bench N=2 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_2gpu_r2final.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'allreduce',d.get('allreduce',{}).get('ms'))
print({k:v for k,v in (d.get('exchange_check') or {}).items() if k not in ('what','rel_err_vs_1gpu_sum_per_segment')})
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 | cut -c1-400
