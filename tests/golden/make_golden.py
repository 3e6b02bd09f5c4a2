"""Generates tests/golden/*.npz by running the UNMODIFIED reference (isls at HEAD, /root/reference) through
oracle/ref_shim.py on small seeded problems.  Container-only (the reference does not travel to the GPU box);
the produced fixtures are committed and are what pins the oracle (oracle/restated.py) and the CUDA path.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models as M, problems as P, ref_shim as S   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _Qs(p):
    return np.stack([np.diag(q) for q in p["Qdiag"]])


def _clip(lo, hi):
    lo, hi = lo.flatten(), hi.flatten()
    return lambda z: np.clip(z, lo, hi)


def _pad(logs):
    n = max(len(l) for l in logs)
    out = np.full((len(logs), n), np.nan)
    for i, l in enumerate(logs):
        out[i, :len(l)] = l
    return out


def golden_ilqr_admm(p, name, rho_u_scalar):
    model = M.make_model(p["model"], dt=p["dt"])
    xs, us, logs, admm_logs = [], [], [], []
    for b in range(p["x0"].shape[0]):
        s = S.make_isls(model, p["N"], p["zs"], _Qs(p), p["seq"], p["u_std"])
        S.init_nominal(s, p["x0"][b], p["u0"])
        kw = dict(project_u=_clip(p["lo_u"], p["hi_u"]), rho_u=rho_u_scalar)
        if p["lo_x"] is not None:
            kw.update(project_x=_clip(p["lo_x"], p["hi_x"]), rho_x=np.stack([np.diag(r) for r in p["rho_x"]]))
        r = S.run_ilqr_admm(s, model, max_iter=p["I_o"], max_admm_iter=p["I_a"],
                            max_line_search_iter=p["L"], tol=p["tol"], **kw)
        xs.append(r["x"]); us.append(r["u"]); logs.append(r["cost_log"])
        al = r["admm_log"].reshape(-1, 2)
        admm_logs.append(al)
        print(name, b, len(r["cost_log"]), r["cost_log"][-1])
    n = max(len(a) for a in admm_logs)
    last_admm = np.full((len(admm_logs), n, 2), np.nan)
    for i, a in enumerate(admm_logs):
        last_admm[i, :len(a)] = a
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x0=p["x0"], x=np.stack(xs), u=np.stack(us),
                        cost_log=_pad(logs), last_admm_log=last_admm)


def golden_ilqr_dp(p, name, max_iter, L):
    model = M.make_model(p["model"], dt=p["dt"])
    xs, us, logs = [], [], []
    for b in range(p["x0"].shape[0]):
        s = S.make_isls(model, p["N"], p["zs"], _Qs(p), p["seq"], p["u_std"])
        S.init_nominal(s, p["x0"][b], p["u0"])
        r = S.run_ilqr_dp(s, model, max_iter=max_iter, max_line_search_iter=L)
        xs.append(r["x"]); us.append(r["u"]); logs.append(r["cost_log"])
        print(name, b, len(r["cost_log"]), r["cost_log"][-1])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x0=p["x0"], x=np.stack(xs), u=np.stack(us),
                        cost_log=_pad(logs), max_iter=max_iter, L=L)


def golden_backward_pass(name, model_name, N, seed):
    """Teacher-forced Riccati pass: random nominal trajectory -> reference get_AB-independent inputs
    (A,B,c,C as arrays) -> reference backward_pass_DP (isls.py:229-308) K,k."""
    rng = np.random.default_rng(seed)
    model = M.make_model(model_name, dt=0.1 if model_name == "car" else 0.01)
    n, m = model.n, model.m
    x = rng.normal(0, 1, (N, n)); u = rng.normal(0, 1, (N, m))
    A, B = model.get_AB(x, u)
    W = rng.normal(0, 1, (N, n + m, n + m))
    C = W @ np.swapaxes(W, -1, -2) + 0.5 * np.eye(n + m)       # SPD with a non-zero Cux block
    c = rng.normal(0, 1, (N, n + m))
    pkg, _ = S.load()
    with S.quiet():
        s = pkg.iSLS(n, m, N)
        s.A, s.B = A, B
        K, k = s.backward_pass_DP(Cts=C, cts=c)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), A=A, B=B, C=C, c=c, K=K, k=k)
    print(name, np.abs(K).max())


def golden_lqt_admm_dp(p, name):
    pkg, _ = S.load()
    N, n, m = p["N"], p["n"], p["m"]
    model = M.make_model("double_integrator", nb_dim=m, dt=p["dt"])
    xs, us, Ks, ks, its, res = [], [], [], [], [], []
    for b in range(p["x0"].shape[0]):
        with S.quiet():
            s = pkg.SLS(n, m, N)
            s.AB = [model.A, model.B]
            s.set_quadratic_cost(p["zs"], _Qs(p), p["seq"], p["u_std"])
            r = s.ADMM_LQT_DP(p["x0"][b], project_x=_clip(p["lo_x"], p["hi_x"]),
                              project_u=_clip(p["lo_u"], p["hi_u"]), max_iter=p["I_a"],
                              rho_x=np.stack([np.diag(q) for q in p["rho_x"]]), rho_u=float(p["rho_u"][0, 0]),
                              tol=p["tol"], log=True)
        x, u, K, k, logs = r
        xs.append(x.reshape(N, n)); us.append(u.reshape(N, m)); Ks.append(K); ks.append(k)
        its.append(len(logs)); res.append(np.array(logs)[-1])
        print(name, b, len(logs))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x0=p["x0"], x=np.stack(xs), u=np.stack(us),
                        K=np.stack(Ks), k=np.stack(ks), iters=np.array(its), last_res=np.stack(res))


def golden_notebook_pins():
    """Known answers printed in the reference's notebooks, re-derived here from HEAD."""
    pkg, _ = S.load()
    from isls.utils import get_double_integrator_AB
    out = {}
    with S.quiet():
        # Double integrator/LQR and SLS with control bounds.ipynb cells 3-8, 11
        n, m, N = 2, 1, 100
        s = pkg.SLS(n, m, N)
        A, B = get_double_integrator_AB(1, nb_deriv=2, dt=0.01)
        s.AB = [A, B]
        zs = np.stack([np.zeros(2), np.array([1.0, 0.0])])
        Qs = np.stack([np.zeros((2, 2)), np.eye(2) * 1e6])
        seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
        s.set_quadratic_cost(zs, Qs, seq, 1e-2)
        x, u = s.solve(np.zeros(2), method="batch")
        out["di_lqt_max_u"] = np.max(np.abs(u))
        out["di_lqt_last_pos"] = x.reshape(N, n)[-1, 0]
        r = s.ADMM_LQT_Batch(np.zeros(2), project_u=lambda z: np.clip(z, -5, 5), rho_u=1e-2, tol=1e-4,
                             max_iter=100, log=True)
        out["di_admm_batch_iters"] = len(r[-1])
        out["di_admm_batch_max_u"] = np.max(r[1])
        r = s.ADMM_LQT_DP(np.zeros(2), project_u=lambda z: np.clip(z, -5, 5), rho_u=1e-1, tol=1e-4,
                          max_iter=2000, log=True)
        out["di_admm_dp_iters"] = len(r[-1])
        out["di_admm_dp_x"] = r[0].reshape(N, n)
        out["di_admm_dp_u"] = r[1].reshape(N, m)
    print({k: v for k, v in out.items() if np.ndim(v) == 0})
    np.savez_compressed(os.path.join(OUT, "notebook_pins.npz"), **out)


if __name__ == "__main__":
    assert S.available(), "needs the reference tree"
    golden_ilqr_admm(P.car_batch(6), "car_ilqr_admm", 10.0)
    golden_ilqr_admm(P.car_batch(3, stress=True), "car_stress_ilqr_admm", 10.0)
    golden_ilqr_admm(P.arm_batch(3), "arm_ilqr_admm", 1e-3)
    golden_ilqr_dp(P.car_batch(4), "car_ilqr_dp", 30, 25)
    golden_ilqr_dp(P.arm_batch(2), "arm_ilqr_dp", 20, 25)
    golden_backward_pass("car_backward_pass", "car", 100, 11)
    golden_backward_pass("arm_backward_pass", "arm3", 100, 12)
    pd = P.di_batch(3)
    golden_lqt_admm_dp(pd, "di_lqt_admm_dp")
    golden_notebook_pins()
