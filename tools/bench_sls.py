#!/usr/bin/env python
"""C4 (SLS-ADMM, double integrator n=4 m=2 N=50, 1,024 problems): CUDA-event times of ADMM_SLS and controller through
the public API, for the ncu captures of k_sls_admm / k_sls_ctrl_* / k_dgemm.  python tools/bench_sls.py [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from scipy.stats import norm
from isls_b200 import SLS, SetConvexSOC, get_double_integrator_AB

Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = 50
rng = np.random.default_rng(1238)
tg = rng.uniform(0.6, 1.0, (Bn, 2))
s = SLS(4, 2, N, batch=Bn)
s.AB = get_double_integrator_AB(2, 2, 1.0 / N)
zs = np.zeros((Bn, 2, 4)); zs[:, 1, :2] = tg
seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
s.set_quadratic_cost(zs, np.stack([np.zeros((4, 4)), np.eye(4) * 1e6]), seq, 1e-2)
s.solve_sls()
mu = np.zeros(3); mu[0] = 1.0
psi = norm.ppf(0.95)
Au = np.diag(np.sqrt(np.array([0.0, 0.01, 0.01])))
proj = SetConvexSOC([np.concatenate([Au, (-mu / psi)[None]], 0), np.concatenate([Au, (mu / psi)[None]], 0)],
                    [np.append(np.zeros(3), 5.0 / psi)] * 2, rho=1e1, max_iter=100, threshold=1e-3)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


h = {}
ms_admm = timed(lambda: h.update(r=s.ADMM_SLS(project_u=proj, max_iter=50, rho_u=1e2, alpha=1.0, tol=1e-3)))
du, phi = h["r"]
ms_ctl = timed(lambda: s.controller(phi, du))
print(json.dumps(dict(B=Bn, admm_ms=round(ms_admm, 3), controller_ms=round(ms_ctl, 3),
                      problems_per_s=round(Bn / (ms_admm + ms_ctl) * 1e3), mean_iters=float(s.last.iters.double().mean()),
                      mean_inner_total=float(s.last.inner_total.double().mean()))))
