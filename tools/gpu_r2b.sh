#!/bin/bash
# round 2, second call: all GPU tests (no -x), bench, projection FMA experiment
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_depthsort.py -q -x --timeout 120 > gpurun_out/pytest_ds_$TAG.log 2>&1; echo "depthsort pytest rc=$?"; tail -5 gpurun_out/pytest_ds_$TAG.log
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 900 --deselect tests/test_gpu_depthsort.py > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "config\[|passed|failed|FAILED|Error|flips|rel errors|history|bit-equal" gpurun_out/pytest_$TAG.log | tail -40
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
timeout 300 python tools/exp_project.py > gpurun_out/exp_project_default_$TAG.json 2> gpurun_out/exp_project_err.log; cat gpurun_out/exp_project_default_$TAG.json
GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_nofmad.so timeout 300 python tools/exp_project.py > gpurun_out/exp_project_nofmad_$TAG.json 2>> gpurun_out/exp_project_err.log; cat gpurun_out/exp_project_nofmad_$TAG.json
