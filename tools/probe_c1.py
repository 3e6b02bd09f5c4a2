import sys, time
sys.path[:0]=["/root/repo","/root/repo/ilqr-admm_b200","/root/repo/tests"]
import torch, numpy as np
from isls_b200 import Bound, SLS, configs, get_double_integrator_AB, solver as S
for B in (1, 32, 1024):
    p = configs.di_batch(B)
    s = SLS(4, 2, p["N"], batch=B)
    s.AB = get_double_integrator_AB(2, 2, p["dt"])
    s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    run = lambda: s.ADMM_LQT_DP(p["x0"], project_x=Bound(p["lo_x"], p["hi_x"]), project_u=Bound(p["lo_u"], p["hi_u"]),
                                rho_x=p["rho_x"], rho_u=p["rho_u"], max_iter=p["I_a"], tol=p["tol"])
    run(); torch.cuda.synchronize()
    for rep in range(3):
        S.profile_enable(True)
        t0 = time.perf_counter(); run(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        prof = S.profile_collect(); S.profile_enable(False)
        it = s.last.admm_iters.cpu().numpy()
        print("B", B, "host enqueue ms %.1f  total ms %.1f" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3), {k: round(v[0], 2) for k, v in prof.items()}, "iters max", it.max(), "mean", it.mean())
