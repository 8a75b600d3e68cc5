#!/bin/bash
set -u
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3
timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | grep -E "lib:|raster|tile_consumed"
for so in build/variants/*.so; do
  [ -f "$so" ] && GSPLAT_B200_LIB=$PWD/$so timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | grep -E "lib:|raster"
done
