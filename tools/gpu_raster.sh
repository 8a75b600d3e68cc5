#!/bin/bash
set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/raster_bench.py 1000000 8 2>&1 | tail -3
for so in build/variants/*.so; do
  [ -f "$so" ] && GSPLAT_B200_LIB=$PWD/$so python tools/raster_bench.py 1000000 8 2>&1 | tail -3
done
