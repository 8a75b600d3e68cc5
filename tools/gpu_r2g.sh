#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 > gpurun_out/pytest_r2g.log 2>&1; echo "rc=$?"
grep -E "^(FAILED|ERROR)|Error|error" gpurun_out/pytest_r2g.log | head -10
T=$(grep -E "^FAILED" gpurun_out/pytest_r2g.log | head -1 | sed 's/FAILED //; s/ - .*//')
echo "first failing: $T"
if [ -n "$T" ]; then timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "$T" -x -q --timeout 500 > gpurun_out/sanitizer_r2g.log 2>&1; grep -E "Invalid|at 0x|by thread|Address|========= .*kernel|in gs::" gpurun_out/sanitizer_r2g.log | head -30; fi
