#!/usr/bin/env python
"""Fixed-budget C5-style car solves at small per-GPU batch sizes (the strong-scaling shards of BASELINE configs[4]):
ms per solve, solves/s and per-kernel-class CUDA-event times.  python tools/bench_small.py [B ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ilqr-admm_b200")):
    sys.path.insert(0, p)
import torch
from isls_b200 import configs, solver as S


def run(B, model="car", reps=3):
    p = configs.car_batch(B, tol=1e-3, I_o=20, I_a=5, L=20) if model == "car" else configs.arm_batch(B)
    kw = dict(rho_u=p["rho_u"], lo_u=p["lo_u"], hi_u=p["hi_u"])
    if p.get("lo_x") is not None:
        kw.update(rho_x=p["rho_x"], lo_x=p["lo_x"], hi_x=p["hi_x"])
    plan = S.Plan(p["model"], p["N"], p["n"], p["m"], p["dt"], p["Qdiag"], p["seq"], p["u_std"], p["L"], **kw)
    sv = S.BatchSolver(plan, B, "cuda:0", max_outer=p["I_o"], max_admm=p["I_a"], logs=False)
    sv.set_inputs(p["x0"], p["u0"], p["zs"])
    for _ in range(2):
        sv.ilqr_admm(tol=p["tol"], fixed_budget=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sv.ilqr_admm(tol=p["tol"], fixed_budget=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    S.profile_enable(True)
    sv.ilqr_admm(tol=p["tol"], fixed_budget=True)
    prof = S.profile_collect()
    S.profile_enable(False)
    return dict(model=model, B=B, ms=round(ms, 3), solves_per_s=round(B / ms * 1e3),
                us_per_launch={k: round(1e3 * v[0] / v[1], 1) for k, v in prof.items()},
                cost_mean=float(sv.out.cost.mean()), u_sum=float(sv.out.u.double().sum()))


if __name__ == "__main__":
    model = os.environ.get("BENCH_MODEL", "car")
    Bs = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384, 32768, 65536]
    for B in Bs:
        print(json.dumps(run(B, model)), flush=True)
