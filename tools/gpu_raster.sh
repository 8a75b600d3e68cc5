#!/bin/bash
set -u
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | grep -E "^E|FAILED|passed|failed" | head -8
timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | grep -E "lib:|raster"
