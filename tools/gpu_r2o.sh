#!/bin/bash
# tiles-per-CTA variants of the compositing kernels + scheduler spread of one-warp CTAs
set -u
mkdir -p gpurun_out
timeout 120 build/ubench_raster > gpurun_out/ubench_raster_where.json; cat gpurun_out/ubench_raster_where.json
for v in default w2 w4; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 300 python tools/raster_bench.py 1000000 10 2>/dev/null | tail -4
done
