#!/bin/bash
set -u
for o in ranges previous; do
GS_FWD_ORDER=$o timeout 120 python tools/raster_bench.py 1000000 8 2>&1 | grep -E "lib:|raster"
GS_FWD_ORDER=$o timeout 200 python tools/run_configs.py 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$o', [(r['views'], round(r['frames_per_s'],1)) for r in d['config3_multiview_batch_1gpu']])"
done
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
