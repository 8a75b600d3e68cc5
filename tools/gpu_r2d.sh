#!/bin/bash
set -u
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_depthsort.py -q -x --timeout 120 2>&1 | tail -3
for v in default st; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 200 python tools/exp_sort.py 2>&1 | tail -1
done
export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_st.so; timeout 200 python tools/exp_sort.py 3000000 2>&1 | tail -1
unset GSPLAT_B200_LIB
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_training.py -m gpu -q -s --timeout 900 > gpurun_out/pytest_full_$TAG.log 2>&1; echo "fullsize pytest rc=$?"
grep -E "config\[|passed|failed|FAILED|Error|assert" gpurun_out/pytest_full_$TAG.log | tail -30
