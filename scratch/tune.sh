for f in 0 2 1 3; do for c in 0 42; do echo "FUSE=$f CPT=$c"; ISLS_FUSE=$f ISLS_LS_CPT=$c python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']), 'ms/step', round(d['ms_per_step'],1), {k:v['ms_per_launch'] for k,v in d['kernels'].items()})"; done; done
