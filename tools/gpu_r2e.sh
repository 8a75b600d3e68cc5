#!/bin/bash
set -u
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_depthsort.py -q -x --timeout 120 2>&1 | tail -3
for v in default st; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 200 python tools/exp_sort.py 2>&1 | tail -1
done
export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_st.so; timeout 200 python tools/exp_sort.py 3000000 2>&1 | tail -1
unset GSPLAT_B200_LIB
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_depthsort.py > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_$TAG.log | tail -20
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d['gpu_launches'])
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
