#!/bin/bash
set -u
TAG=${1:-r2f}
mkdir -p gpurun_out
for v in default st; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 200 python tools/exp_sort.py 2>&1 | tail -1
done
unset GSPLAT_B200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_depthsort.py -q --timeout 300 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q --timeout 900 -k "config1_whole_frame_c0 or config2" 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
print({k:round(v['frac_of_hbm_peak'],3) for k,v in d['kernels'].items()})
PY
