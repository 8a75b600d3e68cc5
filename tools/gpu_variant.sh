#!/bin/bash
# gpu tests and compositing-kernel times of a library variant (build/variants/libgsplat_b200_$1.so) next to the default build
set -u
V=${1:-raw}
mkdir -p gpurun_out
export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$V.so
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_var_$V.log 2>&1; echo "pytest ($V) rc=$?"
tail -2 gpurun_out/pytest_var_$V.log
timeout 300 python tools/raster_bench.py 1000000 10 2>/dev/null | tail -4
unset GSPLAT_B200_LIB
timeout 300 python tools/raster_bench.py 1000000 10 2>/dev/null | tail -4
