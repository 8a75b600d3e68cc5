#!/bin/bash
# gpu tests + bench + inner-loop ceiling
set -u
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
timeout 120 build/ubench_raster > gpurun_out/ubench_raster_$TAG.json; cat gpurun_out/ubench_raster_$TAG.json
