import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT,'ilqr-admm_b200')): sys.path.insert(0,p)
import torch, numpy as np
from isls_b200 import Bound, SLS, configs, get_double_integrator_AB
for B in (1, 32, 1024, 1024, 8192):
    p = configs.di_batch(B)
    s = SLS(4, 2, p["N"], batch=B)
    s.AB = get_double_integrator_AB(2, 2, p["dt"])
    s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    ts=[]
    for r in range(5):
        torch.cuda.synchronize(); t0=time.perf_counter()
        s.ADMM_LQT_DP(p["x0"], project_x=Bound(p["lo_x"], p["hi_x"]), project_u=Bound(p["lo_u"], p["hi_u"]), rho_x=p["rho_x"], rho_u=p["rho_u"], max_iter=p["I_a"], tol=p["tol"])
        torch.cuda.synchronize(); ts.append(round((time.perf_counter()-t0)*1e3,1))
    print(B, ts, float(s.last.admm_iters.double().mean()), int(s.last.admm_iters.max()))
