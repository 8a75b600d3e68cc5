#!/bin/bash
set -u
for n in 250000 500000 1000000; do
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'scatter_kernel|chunk_walk' --launch-skip 6 --launch-count 2 --csv python tools/bin_bench.py $n 2 2>/dev/null | grep -E "scatter_kernel|chunk_walk|^algo" | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-40,100-260
timeout 100 python tools/bin_bench.py $n 6 2>&1 | grep "algo 1"
done
