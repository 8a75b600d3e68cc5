#!/bin/bash
set -u
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | tail -4
