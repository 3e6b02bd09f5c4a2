#!/usr/bin/env python
"""Helper of tests/test_gpu_kernel_variants.py: solves two small problems through the public API in THIS process (whose
environment selects the kernel variants: ISLS_FF_STAGES, ISLS_ADMM_STAGES are read once per process by the library)
and saves every output array."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import gpu_util as G
from isls_b200 import configs

out = {}
for tag, p in (("arm", configs.arm_batch(40, I_o=4, I_a=4)),                      # state + control bounds: k_admm
               ("park", configs.parking_batch(6, N=120, dt=0.125, I_o=3, I_a=4)),     # obstacle sets: k_admm + k_obst_project
               ("car", configs.car_batch(70, I_o=3, I_a=3))):                       # control bounds only (fused update)
    o = G.run_ilqr_admm(p)
    for k, v in o.items():
        out[tag + "_" + k] = v
o = G.run_isls_admm(configs.arm_robust_batch(3, I_o=3, I_a=3))      # robust iSLS-ADMM: k_isls_cols
for k, v in o.items():
    out["robust_" + k] = v
o = G.run_lqt_admm_dp(configs.di_batch(5, max_iter=60))                 # LQT-ADMM: k_lqt_admm<., SM>
for k, v in o.items():
    out["lqt_" + k] = v
o = G.run_lqt_admm_dp(configs.di_obstacle_batch(2, max_iter=12))        # ... with the spherical-obstacle projection
for k, v in o.items():
    out["lqtobs_" + k] = v
# SLS-ADMM + controller (batched shared-column controller vs the per-problem dense form, ISLS_SLS_CTRL_DENSE)
from scipy.stats import norm                                                              # noqa: E402
from isls_b200 import SLS, SetConvexSOC, get_double_integrator_AB                          # noqa: E402
N_ = 30
s_ = SLS(4, 2, N_, batch=5)
s_.AB = get_double_integrator_AB(2, 2, 1.0 / N_)
zs_ = np.zeros((5, 2, 4)); zs_[:, 1, :2] = np.random.default_rng(3).uniform(0.6, 1.0, (5, 2))
seq_ = np.zeros(N_, dtype=np.int32); seq_[-1] = 1
s_.set_quadratic_cost(zs_, np.stack([np.zeros((4, 4)), np.eye(4) * 1e6]), seq_, 1e-2)
mu_ = np.zeros(3); mu_[0] = 1.0
psi_ = norm.ppf(0.95)
Au_ = np.diag(np.sqrt(np.array([0.0, 0.01, 0.01])))
proj_ = SetConvexSOC([np.concatenate([Au_, (-mu_ / psi_)[None]], 0), np.concatenate([Au_, (mu_ / psi_)[None]], 0)],
                     [np.append(np.zeros(3), 5.0 / psi_)] * 2, rho=1e1, max_iter=100, threshold=1e-3)
du_, phi_ = s_.ADMM_SLS(project_u=proj_, max_iter=20, rho_u=1e2, alpha=1.0, tol=1e-3, fixed_budget=True)
K_, k_ = s_.controller(phi_, du_)
out.update(sls_du=du_.cpu().numpy(), sls_phi=phi_.cpu().numpy(), sls_K=K_.cpu().numpy(), sls_k=k_.cpu().numpy())
np.savez(sys.argv[1], **out)
print("variant ok", sorted(k for k in os.environ if k.startswith("ISLS_")))
