for st in 0 -1; do echo "FF_STAGES=$st"; ISLS_FF_STAGES=$st python tools/bench_configs.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items(): print(k, {a:b for a,b in v.items()})"; done
python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']), 'ms/step', round(d['ms_per_step'],1), {k:v['ms_per_launch'] for k,v in d['kernels'].items()})"
