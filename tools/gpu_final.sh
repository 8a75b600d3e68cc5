#!/bin/bash
# end-of-round check: all gpu tests, the bench with the driver's flags, the ncu launch list of the same build
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_${TAG}_err.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',d['e2e'], {k:round(v['ms'],4) for k,v in d['kernels'].items()})
print(d['roofline']['frac'], d['roofline']['issue'].get('frac_of_issue_slot_roof'), d['cpu_baseline']['value'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
