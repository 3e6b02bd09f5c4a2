#!/bin/bash
# GPU-side sweep of the line-search CTA shapes (ISLS_LS_CPT knob of csrc/isls_b200.cu); prints ms per launch per kernel.
for v in 0 5 4; do
  echo "ISLS_LS_CPT=$v"
  ISLS_LS_CPT=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  solves/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), {k:v['ms_per_launch'] for k,v in d['kernels'].items()})"
done
