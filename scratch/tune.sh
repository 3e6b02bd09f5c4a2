for c in 0 73 74 75; do echo "CPT=$c"; ISLS_LS_CPT=$c python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']), 'ms/step', round(d['ms_per_step'],1), {k:v['ms_per_launch'] for k,v in d['kernels'].items() if k=='linesearch'})"; done
