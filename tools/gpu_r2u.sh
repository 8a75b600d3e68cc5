#!/bin/bash
set -u
mkdir -p gpurun_out
for v in default tt512 tt1024; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  echo "== $v"; timeout 300 python tools/bin_bench.py 1000000 10 2>/dev/null | grep "algo 1"
done
unset GSPLAT_B200_LIB
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u_err.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2u.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'weighted_sum|tile_tables' -c 6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "weighted_sum|tile_tables|gpu__time_duration" | head -20
