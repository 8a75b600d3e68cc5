#!/bin/bash
set -u
TAG=${1:-b}
python tools/bin_bench.py 1000000 10 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:'chunk_walk|scatter_kernel|column_prefix|tile_scan|super_prefix' --launch-skip 8 --launch-count 4 \
    -o gpurun_out/prof_bin_$TAG -f python tools/bin_bench.py 1000000 2 > gpurun_out/ncu_bin_$TAG.log 2>&1; echo "ncu rc=$?"
