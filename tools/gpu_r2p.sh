#!/bin/bash
set -u
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
