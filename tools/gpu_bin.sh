#!/bin/bash
set -u
TAG=${1:-b}
timeout 120 python tools/bin_bench.py 1000000 10 2>&1 | tail -6
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 100 -k "sort_keys or cuda_matches or optimistic or full_size or golden" 2>&1 | tail -5
timeout 200 ncu --set full --clock-control none -k regex:'coarse_emit|piece_map|fine_count|block_scan|fine_write|tile_scan|ranges_kernel|Onesweep|Histogram' --launch-skip 0 --launch-count 40 \
    -o gpurun_out/prof_bin_$TAG -f python tools/bin_bench.py 1000000 1 > gpurun_out/ncu_bin_$TAG.log 2>&1; echo "ncu rc=$?"
