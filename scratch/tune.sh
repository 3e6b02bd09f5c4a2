run() { python bench.py --steps 3 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']), 'ms/step', round(d['ms_per_step'],1))"; }
for ch in 1 3 4 6 8; do echo "CHUNKS=$ch"; ISLS_CHUNKS=$ch run; done
