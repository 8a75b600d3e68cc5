#!/bin/bash
set -u
timeout 100 python tools/bin_bench.py 1000000 8 2>&1 | grep "algo 1"
for so in build/variants/*.so; do
  echo $so; GSPLAT_B200_LIB=$PWD/$so timeout 100 python tools/bin_bench.py 1000000 8 2>&1 | grep "algo 1"
done
