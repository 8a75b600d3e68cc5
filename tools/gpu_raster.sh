#!/bin/bash
# gpu tests with the default build (per-test and overall timeouts: a hung kernel must not eat the GPU budget),
# then the raster micro-bench for the default build and every variant
set -u
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | tail -3
for so in build/variants/*.so; do
  [ -f "$so" ] && GSPLAT_B200_LIB=$PWD/$so timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | tail -3
done
