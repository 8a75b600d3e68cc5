#!/bin/bash
set -u
TAG=${1:-q2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_$TAG.log
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1), {k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
done
