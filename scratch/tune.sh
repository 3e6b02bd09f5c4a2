for c in 0 1 2 3; do echo "COMPACT=$c"; ISLS_COMPACT=$c python bench.py --steps 2 --warmup 2 --no-cpu-baseline --early-exit | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('early-exit value',round(d['value']), 'ms/step', round(d['ms_per_step'],1), {k:v['ms_total'] for k,v in d['kernels'].items()})"; done
for c in 0 1 2; do echo "COMPACT=$c"; ISLS_COMPACT=$c python tools/bench_configs.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if k[:2] in ('C2','C3'): print(k, {a:b for a,b in v.items()})"; done
