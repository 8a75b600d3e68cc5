#!/bin/bash
# round 2, first call: all GPU tests (with the new whole-frame ones) + a bench line
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
nproc; nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "config\[|passed|failed|error|flips|rel errors|history" gpurun_out/pytest_$TAG.log | tail -40
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'cpu', d.get('cpu_baseline'))
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
