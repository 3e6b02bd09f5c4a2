#!/usr/bin/env python
"""Tiny run of every kernel family (for compute-sanitizer memcheck / quick smoke on a GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gpu_util as G
from isls_b200 import SetConvexSOC, configs, solver as S

p = configs.car_batch(40, I_o=2, I_a=2, L=20)
o = G.run_ilqr_admm(p)
print("car ilqr_admm", o["cost"][:3])
o = G.run_ilqr_admm(p, fixed_budget=True)
p = configs.car_batch(4100, I_o=2, I_a=2, L=20)        # > 1536 tiles? no: 129 tiles -> staged ff path
o = G.run_ilqr_admm(p, want_masks=False)
p = configs.arm_batch(33, I_o=2, I_a=2, L=5)
o = G.run_ilqr_admm(p)
print("arm ilqr_admm", o["cost"][:3])
o = G.run_ilqr_dp(configs.car_batch(35), 3, 25)
print("car ilqr", o["cost"][:3])
o = G.run_ilqr_dp(configs.arm_batch(5), 2, 25)
pd = configs.di_batch(5, max_iter=30)
o = G.run_lqt_admm_dp(pd)
print("lqt", o["admm_iters"][:, 0])
# stage-level entry points
g = np.load(os.path.join(ROOT, "tests", "golden", "car_backward_pass.npz"))
t = lambda a: torch.as_tensor(a, device="cuda:0")[None].repeat(3, *([1] * a.ndim))
K, k, bad = S.riccati(t(g["A"]), t(g["B"]), t(g["c"]), t(g["C"]))
x = torch.randn(3, 77, device="cuda:0", dtype=torch.float64)
z, lam = torch.zeros_like(x), torch.zeros_like(x)
S.admm_project_dual(x, z, lam, torch.full((77,), -0.5, dtype=torch.float64), torch.full((77,), 0.5, dtype=torch.float64),
                    want_mask=True)
# SLS path
import test_gpu_sls as T
from isls_b200 import get_double_integrator_AB
A, B = get_double_integrator_AB(2, 2, 0.05)
s = T._make_sls(4, 2, 20, A, B, np.array([[0.8, 0.7], [0.6, 0.9], [1.0, 1.0]]))
PHI, du = s.solve_sls()
from scipy.stats import norm
mu = np.zeros(3); mu[0] = 1.0
psi = norm.ppf(0.95)
Au = np.diag(np.sqrt(np.array([0.0, 0.01, 0.01])))
A_ = [np.concatenate([Au, (-mu / psi)[None]], 0), np.concatenate([Au, (mu / psi)[None]], 0)]
b_ = [np.append(np.zeros(3), 5.0 / psi)] * 2
du, phi = s.ADMM_SLS(project_u=SetConvexSOC(A_, b_, rho=1e1, max_iter=50, threshold=1e-3), max_iter=10, rho_u=1e2)
K, k = s.controller(phi, du)
torch.cuda.synchronize()
print("sls", s.last.iters.cpu().numpy(), float(K.abs().max()))
print("SANITY OK")
