#!/usr/bin/env python
"""Times the BASELINE.json configs C1-C4 on one GPU (reference stop rules unless --fixed): problems/s per config.
Not the headline bench (bench.py = C5); used for profiles/README.md."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from isls_b200 import Bound, SetConvexSOC, SLS, configs, get_double_integrator_AB, iSLS, solver as S


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def isls_case(p, fixed):
    B = p["x0"].shape[0]
    s = iSLS(p["n"], p["m"], p["N"], batch=B)
    s.forward_model = (p["model"], {"dt": p["dt"]})
    s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
    kw = {}
    if p.get("lo_x") is not None:
        kw.update(project_x=Bound(p["lo_x"], p["hi_x"]), rho_x=p["rho_x"])
    kw.update(project_u=Bound(p["lo_u"], p["hi_u"]), rho_u=p["rho_u"])

    def run():
        s.set_initial(p["x0"], p["u0"])
        return s.ilqr_admm(max_iter=p["I_o"], max_admm_iter=p["I_a"], max_line_search_iter=p["L"], tol=p["tol"],
                           fixed_budget=fixed, **kw)
    ms = timed(run)
    out = run()
    S.profile_enable(True)
    run()
    prof = S.profile_collect()
    S.profile_enable(False)
    return ms, out, {k: round(v[0], 3) for k, v in prof.items()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--fixed", action="store_true")
    ap.add_argument("--quick", action="store_true", help="C2 and C3 only")
    a = ap.parse_args()
    res = {}
    # C2: car, 4,096 problems
    ms, out, prof = isls_case(configs.car_batch(4096), a.fixed)
    res["C2 car iLQR-ADMM B=4096"] = dict(ms=round(ms, 2), problems_per_s=round(4096 / ms * 1e3),
                                          mean_outer=float(out.outer_iters.double().mean()), kernels_ms=prof)
    # C3: arm, 16,384 problems
    ms, out, prof = isls_case(configs.arm_batch(16384), a.fixed)
    res["C3 arm iLQR-ADMM B=16384"] = dict(ms=round(ms, 2), problems_per_s=round(16384 / ms * 1e3),
                                           mean_outer=float(out.outer_iters.double().mean()),
                                           mean_cost=float(out.cost.mean()), kernels_ms=prof)
    if a.quick:
        print(json.dumps(res))
        sys.exit(0)
    # C1: LQT-ADMM DP double integrator, B = 1 and 1,024
    for B in (1, 1024):
        p = configs.di_batch(B)
        s = SLS(4, 2, p["N"], batch=B)
        s.AB = get_double_integrator_AB(2, 2, p["dt"])
        s.set_quadratic_cost(p["zs"], p["Qdiag"], p["seq"], p["u_std"])
        run = lambda: s.ADMM_LQT_DP(p["x0"], project_x=Bound(p["lo_x"], p["hi_x"]), project_u=Bound(p["lo_u"], p["hi_u"]),
                                    rho_x=p["rho_x"], rho_u=p["rho_u"], max_iter=p["I_a"], tol=p["tol"])
        ms = timed(run)
        res["C1 LQT-ADMM-DP DI B=%d" % B] = dict(ms=round(ms, 2), problems_per_s=round(B / ms * 1e3),
                                                 mean_iters=float(s.last.admm_iters.double().mean()))
    # C4: SLS-ADMM, 1,024 problems
    from scipy.stats import norm
    Bn, N = 1024, 50
    rng = np.random.default_rng(1238)
    tg = rng.uniform(0.6, 1.0, (Bn, 2))
    A, Bm = get_double_integrator_AB(2, 2, 1.0 / N)
    s = SLS(4, 2, N, batch=Bn)
    s.AB = [A, Bm]
    zs = np.zeros((Bn, 2, 4)); zs[:, 1, :2] = tg
    seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
    t0 = time.perf_counter()
    s.set_quadratic_cost(zs, np.stack([np.zeros((4, 4)), np.eye(4) * 1e6]), seq, 1e-2)
    s.solve_sls(); torch.cuda.synchronize()
    t_plan = (time.perf_counter() - t0) * 1e3
    mu = np.zeros(3); mu[0] = 1.0
    psi = norm.ppf(0.95)
    Au = np.diag(np.sqrt(np.array([0.0, 0.01, 0.01])))
    A_ = [np.concatenate([Au, (-mu / psi)[None]], 0), np.concatenate([Au, (mu / psi)[None]], 0)]
    b_ = [np.append(np.zeros(3), 5.0 / psi)] * 2
    proj = SetConvexSOC(A_, b_, rho=1e1, max_iter=100, threshold=1e-3)
    holder = {}

    def run4():
        holder["r"] = s.ADMM_SLS(project_u=proj, max_iter=50, rho_u=1e2, alpha=1.0, tol=1e-3, fixed_budget=a.fixed)
    ms_admm = timed(run4)
    du, phi = holder["r"]
    ms_ctl = timed(lambda: s.controller(phi, du))
    res["C4 SLS-ADMM DI B=1024"] = dict(plan_ms=round(t_plan, 1), admm_ms=round(ms_admm, 2), controller_ms=round(ms_ctl, 2),
                                        problems_per_s=round(Bn / (ms_admm + ms_ctl) * 1e3),
                                        mean_iters=float(s.last.iters.double().mean()))
    # ---- stage-level generic-operator Riccati pass (isls_riccati_f64, SURVEY 8a a1 / 8d roofline (i)): dense A, B, c, C
    # from HBM in the reference's natural layouts, 608 B per problem-step for the car shape
    for (n_, m_, Bq) in ((4, 2, 65536), (9, 3, 16384)):
        N_ = 100
        g = torch.Generator(device="cuda").manual_seed(1)
        A_ = torch.eye(n_, dtype=torch.float64, device="cuda").expand(Bq, N_, n_, n_).contiguous()
        A_ += 0.05 * torch.randn(Bq, N_, n_, n_, dtype=torch.float64, device="cuda", generator=g)
        B_ = 0.1 * torch.randn(Bq, N_, n_, m_, dtype=torch.float64, device="cuda", generator=g)
        c_ = torch.randn(Bq, N_, n_ + m_, dtype=torch.float64, device="cuda", generator=g)
        W_ = torch.randn(Bq, N_, n_ + m_, n_ + m_, dtype=torch.float64, device="cuda", generator=g)
        C_ = W_ @ W_.transpose(-1, -2) + torch.eye(n_ + m_, dtype=torch.float64, device="cuda")
        ms = timed(lambda: S.riccati(A_, B_, c_, C_))
        byts = Bq * N_ * 8.0 * (n_ * n_ + n_ * m_ + (n_ + m_) + (n_ + m_) ** 2 + m_ * n_ + m_)
        res["a1 generic-operator Riccati pass n=%d m=%d N=100 B=%d" % (n_, m_, Bq)] = dict(
            ms=round(ms, 3), passes_per_s=round(Bq / ms * 1e3), algorithmic_GBps=round(byts / ms / 1e6, 1),
            frac_of_hbm_peak=round(byts / ms / 1e6 / 6542.7, 3))
        del A_, B_, c_, W_, C_
    # ---- widened rows: same measurement, reference stop rules
    import gpu_util

    def widened(name, p, runner, **kw):
        B = p["x0"].shape[0]
        holder = {}

        def run():
            holder["o"] = runner(p, **kw)
        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3            # includes the host-side staging and the D2H of the results
        S.profile_enable(True)
        run()
        prof = S.profile_collect()
        S.profile_enable(False)
        o = holder["o"]
        res[name] = dict(ms=round(ms, 2), problems_per_s=round(B / ms * 1e3),
                         mean_outer=float(np.mean(o["outer_iters"])), mean_cost=float(np.mean(o["cost"])),
                         kernels_ms={k: round(v[0], 3) for k, v in prof.items()})
    widened("8f#3 Tutorial Tassa car + pseudo-Huber iLQR-ADMM N=150 B=4096", configs.tassa_batch(4096),
            gpu_util.run_ilqr_admm, fixed_budget=a.fixed, want_masks=False)
    widened("8f#3 Tutorial Tassa car + pseudo-Huber iLQR (dp) N=150 B=4096", configs.tassa_batch(4096),
            lambda p, **kw: gpu_util.run_ilqr_dp(p, 100, 40, **kw), fixed_budget=a.fixed)
    widened("8f#2 parking between two cars (obstacle-set state projection) N=200 B=1024",
            configs.parking_batch(1024, N=200, dt=0.075), gpu_util.run_ilqr_admm, fixed_budget=a.fixed)
    widened("8f#1 robust iSLS-ADMM 3-DoF arm N=100 B=1024", configs.arm_robust_batch(1024), gpu_util.run_isls_admm,
            fixed_budget=a.fixed)
    print(json.dumps(res, indent=1))
