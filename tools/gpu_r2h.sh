#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 2>&1 | tail -3
timeout 200 python tools/bin_bench.py 2>&1 | tail -4
bash tools/gpu_multi.sh 2 r2h
