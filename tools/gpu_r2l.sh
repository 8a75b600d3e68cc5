#!/bin/bash
set -u
TAG=${1:-r2l}
mkdir -p gpurun_out
for v in default pv; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 300 python tools/exp_project.py 2>/dev/null | tail -1
done
export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_pv.so
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_training.py -m gpu -q --timeout 600 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q --timeout 900 -k "aniso or config4" 2>&1 | tail -3
