import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'ilqr-admm_b200'); sys.path.insert(0,'tests')
import gpu_util as g
from oracle import problems as P, restated as R
p = P.car_batch(64)
out = g.run_ilqr_dp(p, 30, 25)
o = R.ilqr_dp(p, max_iter=30, L=25)
a,b = out['cost_log'], o['cost_log']
m = ~np.isnan(b)
rel = np.where(m, np.abs(a-b)/np.abs(b), 0)
print('max rel per problem', np.sort(rel.max(1))[-8:])
bi = rel.max(1).argmax()
print('worst problem', bi, 'iters', out['n_log'][bi])
print(np.c_[a[bi], b[bi], rel[bi]][:out['n_log'][bi]])
print('alpha gpu', out['alpha_idx'][bi].ravel()[:30]); print('alpha ora', o['alpha_idx'][bi][:30])
print('K rel', np.abs(out['K']-o['K']).max()/np.abs(o['K']).max())
