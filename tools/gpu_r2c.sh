#!/bin/bash
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -s --timeout 900 > gpurun_out/pytest_full_$TAG.log 2>&1; echo "fullsize pytest rc=$?"
grep -E "config\[|passed|failed|FAILED|Error|assert" gpurun_out/pytest_full_$TAG.log | tail -30
for v in default st; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 200 python tools/exp_sort.py 2>&1 | tail -1
done
export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_st.so; timeout 200 python tools/exp_sort.py 3000000 2>&1 | tail -1
for v in f0b0 f1b0 f2b0 f1b1 f2b2; do
  GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so timeout 200 python tools/raster_bench.py 1000000 10 2>&1 | tail -4
done
