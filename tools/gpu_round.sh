#!/bin/bash
# one gpurun call: gpu tests, bench, timeline, ncu launch list, ncu full capture of the heavy kernels
set -u
TAG=${1:-x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
timeout 200 python tools/timeline.py > gpurun_out/timeline_$TAG.txt 2>&1; echo "timeline rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'raster_|scatter_kernel|chunk_walk|column_prefix|project_|tile_order' --launch-skip 40 --launch-count 12 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
