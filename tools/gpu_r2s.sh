#!/bin/bash
# profiles of record for round 2: bench, launch list, ncu --set full of every kernel of one steady-state frame
set -u
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_|scatter_kernel|chunk_walk|column_prefix|tile_tables|project_|tile_order|depth_sort|weighted_sum' --launch-skip 60 --launch-count 16 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
