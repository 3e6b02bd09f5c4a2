for e in 1 0; do echo "NO_FUSED_UPDATE=$e"; ISLS_NO_FUSED_UPDATE=$e python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']), 'ms/step', round(d['ms_per_step'],1), {k:v['ms_per_launch'] for k,v in d['kernels'].items()})"; done
