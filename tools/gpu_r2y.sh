#!/bin/bash
set -u
for v in default pb5 pb6; do
  if [ $v = default ]; then unset GSPLAT_B200_LIB; else export GSPLAT_B200_LIB=$PWD/build/variants/libgsplat_b200_$v.so; fi
  timeout 300 python tools/exp_project.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['lib'][-30:], d['ref']['ms'], d['aniso']['ms'])"
done
