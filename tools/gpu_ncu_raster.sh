#!/bin/bash
set -u
python tools/raster_bench.py 1000000 4 2>&1 | tail -3
GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_nored.so python tools/raster_bench.py 1000000 4 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:'raster_bwd' --launch-skip 3 --launch-count 1 -o gpurun_out/prof_bwd_def -f python tools/raster_bench.py 1000000 3 > gpurun_out/ncu_bwd_def.log 2>&1; echo rc=$?
GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_nored.so ncu --set full --clock-control none --import-source on -k regex:'raster_bwd' --launch-skip 3 --launch-count 1 -o gpurun_out/prof_bwd_nored -f python tools/raster_bench.py 1000000 3 > gpurun_out/ncu_bwd_nored.log 2>&1; echo rc=$?
