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


def golden_lqt_admm_batch(p, name):
    """SLS.ADMM_LQT_Batch of the unmodified reference (isls/sls.py:250-294) on the C1 problems."""
    pkg, _ = S.load()
    N, n, m = p["N"], p["n"], p["m"]
    model = M.make_model("double_integrator", nb_dim=m, dt=p["dt"])
    xs, us, its, res = [], [], [], []
    for b in range(p["x0"].shape[0]):
        with S.quiet():
            s = pkg.SLS(n, m, N)
            s.AB = [model.A, model.B]
            s.set_quadratic_cost(p["zs"], _Qs(p), p["seq"], p["u_std"])
            x, u, logs = s.ADMM_LQT_Batch(p["x0"][b], project_x=_clip(p["lo_x"], p["hi_x"]),
                                          project_u=_clip(p["lo_u"], p["hi_u"]), max_iter=p["I_a"],
                                          rho_x=np.stack([np.diag(q) for q in p["rho_x"]]),
                                          rho_u=float(p["rho_u"][0, 0]), tol=p["tol"], log=True)
        xs.append(x.reshape(N, n)); us.append(u.reshape(N, m)); its.append(len(logs)); res.append(np.array(logs)[-1])
        print(name, b, len(logs))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x0=p["x0"], x=np.stack(xs), u=np.stack(us),
                        iters=np.array(its), last_res=np.stack(res))


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


def golden_sls(name="sls_admm"):
    """SLS.solve_sls / ADMM_SLS / controller of the unmodified reference on (a) the notebook problem
    'LQR and SLS with control bounds' (n=2, m=1, N=100; known printout: "can't improve anymore at iteration 25",
    residual 8.50e-13 3.77e-01) and (b) a C4-style problem (n=4, m=2, N=50, dt=1/50) for 3 targets."""
    from scipy.stats import norm
    pkg, _ = S.load()
    from isls.utils import get_double_integrator_AB
    from isls.projections import project_set_convex, project_soc_unit
    out = {}

    def cones(pos_dim, upper, lower, var_x0=0.01, p=0.95):
        mu = np.zeros(pos_dim + 1); mu[0] = 1.0
        sigma = np.zeros(pos_dim + 1); sigma[1:] = var_x0
        psi = norm.ppf(p)
        Au = np.diag(np.sqrt(sigma)); bu = np.zeros(pos_dim + 1)
        A_ = [np.concatenate([Au, (-mu / psi)[None]], 0), np.concatenate([Au, (mu / psi)[None]], 0)]
        b_ = [np.append(bu, upper / psi), np.append(bu, -lower / psi)]
        return A_, b_

    def run(tag, pos_dim, N, dt, Qf, u_std, targets, bound):
        n, m = 2 * pos_dim, pos_dim
        A, B = get_double_integrator_AB(pos_dim, nb_deriv=2, dt=dt)
        A_, b_ = cones(pos_dim, bound, -bound)
        proj = lambda y: project_set_convex(y, A_, b_, projections=[project_soc_unit] * 2, rho=1e1, max_iter=100,
                                            threshold=1e-3)
        res = dict(du0=[], du=[], phic=[], iters=[], last=[], logs=[], u_cl=[])
        for tg in targets:
            with S.quiet():
                sls = pkg.SLS(n, m, N); sls.AB = [A, B]
                zs = np.stack([np.zeros(n), np.concatenate([tg, np.zeros(pos_dim)])])
                Qs = np.stack([np.zeros((n, n)), np.eye(n) * Qf])
                seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
                sls.set_quadratic_cost(zs, Qs, seq, u_std)
                PHI0, du0 = sls.solve_sls()
                du, PHI, log = sls.ADMM_SLS(project_u=proj, max_iter=50, rho_u=1e2, alpha=1., tol=1e-3, log=True)
                K, k = sls.controller(PHI, du)
                # closed-loop controls for a few initial positions (well-conditioned function of the controller)
                x0s = np.zeros((3, n)); x0s[:, :pos_dim] = np.array([[0.05], [-0.1], [0.12]])[:, :1]
                _, u_cl = sls.get_trajectory_sls(x0s, K, k)
            res["du0"].append(du0); res["du"].append(du); res["phic"].append(PHI[:, :pos_dim])
            res["iters"].append(len(log)); res["last"].append(np.array(log)[-1]); res["u_cl"].append(u_cl)
            lg = np.full((50, 2), np.nan); lg[:len(log)] = np.array(log); res["logs"].append(lg)
            print(tag, tg, len(log), np.array(log)[-1])
        out[tag + "_PHI0"] = PHI0
        out[tag + "_A_"] = np.stack(A_); out[tag + "_b_"] = np.stack(b_)
        out[tag + "_targets"] = np.stack(targets)
        for kk, v in res.items():
            out[tag + "_" + kk] = np.stack(v)

    run("nb", 1, 100, 0.01, 1e6, 1e-2, [np.array([1.0])], 5.0)
    run("c4", 2, 50, 1.0 / 50, 1e6, 1e-2, [np.array([1.0, 1.0]), np.array([0.7, 0.9]), np.array([0.6, 0.65])], 5.0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def golden_mc(name="mc_rollouts"):
    """get_trajectory_dp / get_trajectory_sls / get_trajectory_batch of the unmodified reference (noise_scale = 0) for
    (a) SLS with the double integrator (sls_base.py:62-105) and (b) iSLS with the car (isls_base.py:28-71)."""
    pkg, _ = S.load()
    from isls.utils import get_double_integrator_AB
    rng = np.random.default_rng(99)
    out = {}
    with S.quiet():
        n, m, N = 4, 2, 30
        A, B = get_double_integrator_AB(2, nb_deriv=2, dt=0.05)
        sls = pkg.SLS(n, m, N); sls.AB = [A, B]
        K = rng.normal(0, 0.3, (N, m, n)); k = rng.normal(0, 0.5, (N, m))
        Ks = np.tril(np.ones((N, N)))[:, None, :, None] * rng.normal(0, 0.05, (N, m, N, n))
        Ks = Ks.reshape(N * m, N * n); ks = rng.normal(0, 0.5, N * m)
        x0 = rng.normal(0, 1.0, (7, n)); us = rng.normal(0, 1.0, (N, m))
        out.update(di_K=K, di_k=k, di_Ks=Ks, di_ks=ks, di_x0=x0, di_us=us)
        out["di_dp_x"], out["di_dp_u"] = sls.get_trajectory_dp(x0, K, k)
        out["di_sls_x"], out["di_sls_u"] = sls.get_trajectory_sls(x0, Ks, ks)
        out["di_batch_x"], out["di_batch_u"] = sls.get_trajectory_batch(x0, us)
        # iSLS + car: nominal trajectory needed by get_trajectory_sls
        model = M.make_model("car", dt=0.1)
        N2 = 25
        s = pkg.iSLS(4, 2, N2); s.forward_model = model.f
        u_nom = rng.normal(0, 0.2, (N2, 2)); x0n = np.array([0.5, -0.3, 1.0, 0.4])
        x_nom, _ = s.rollout_batch(x0n[None], u_nom[None])
        s.x_nom, s.u_nom = x_nom[0], u_nom
        Kc = rng.normal(0, 0.2, (N2, 2, 4)); kc = rng.normal(0, 0.3, (N2, 2))
        Kcs = (np.tril(np.ones((N2, N2)))[:, None, :, None] * rng.normal(0, 0.03, (N2, 2, N2, 4))).reshape(N2 * 2, N2 * 4)
        kcs = rng.normal(0, 0.1, N2 * 2)
        x0c = x0n + rng.normal(0, 0.2, (6, 4))
        out.update(car_x_nom=x_nom[0], car_u_nom=u_nom, car_K=Kc, car_k=kc, car_Ks=Kcs, car_ks=kcs, car_x0=x0c)
        # NOTE isls_base.py:39-42 / 68-71 return element [0] when x0.ndim == 2: call per sample
        xs, uu = [], []
        for r in x0c:
            a, b = s.get_trajectory_dp(r[None], Kc, kc); xs.append(a); uu.append(b)
        out["car_dp_x"], out["car_dp_u"] = np.stack(xs), np.stack(uu)
        xs, uu = [], []
        for r in x0c:
            a, b = s.get_trajectory_sls(r[None], Kcs, kcs); xs.append(a); uu.append(b)
        out["car_sls_x"], out["car_sls_u"] = np.stack(xs), np.stack(uu)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("_x")})


def golden_tutorial(B=3, N=150):
    """Tutorial problem (Tassa car parking + pseudo-Huber cost): reference solve(method='dp') (cell 20) and
    ilqr_admm with control limits (cell 27), analytic get_AB / get_Cs standing in for autograd (not installed)."""
    p = P.tassa_batch(B, N=N)
    model = M.make_model("tassa_car", dt=p["dt"])
    out = {}
    for tag, run in (("dp", S.run_tutorial_dp), ("admm", S.run_tutorial_admm)):
        xs, us, logs, admm = [], [], [], []
        for b in range(B):
            s = S.make_isls_tutorial(model, p)
            S.init_nominal(s, p["x0"][b], p["u0"])
            r = run(s, model, p)
            xs.append(r["x"]); us.append(r["u"]); logs.append(r["cost_log"])
            if "admm_log" in r:
                admm.append(r["admm_log"].reshape(-1, 2))
            print("tutorial", tag, b, len(r["cost_log"]), r["cost_log"][0], r["cost_log"][-1])
        out.update({"x_" + tag: np.stack(xs), "u_" + tag: np.stack(us), "cost_log_" + tag: _pad(logs)})
        if admm:
            k = max(len(a) for a in admm)
            la = np.full((B, k, 2), np.nan)
            for i, a in enumerate(admm):
                la[i, :len(a)] = a
            out["last_admm_log"] = la
    np.savez_compressed(os.path.join(OUT, "tutorial_tassa.npz"), x0=p["x0"], u0=p["u0"], N=N, **out)


def golden_tutorial_fullsize(N=500):
    """The same problem at the notebooks' own size (Tutorial cell 4 / `Car/Replicate of control-limited ddp car
    example.ipynb` cells 5, 14-15, 19-21: T = 15 s, N = 500, x0 = (1, 1, 3pi/2, 0), max_iter=100 resp. 50 x 5 ADMM
    iterations x 40 candidates, rho_u = diag(1e-1, 1e-2)); the notebooks draw u0 unseeded, here it is seeded."""
    p = P.tassa_batch(1, N=N)
    p["x0"][0] = np.array([1.0, 1.0, 1.5 * np.pi, 0.0])
    model = M.make_model("tassa_car", dt=p["dt"])
    out = {}
    for tag, run in (("dp", S.run_tutorial_dp), ("admm", S.run_tutorial_admm)):
        s = S.make_isls_tutorial(model, p)
        S.init_nominal(s, p["x0"][0], p["u0"])
        r = run(s, model, p)
        print("tutorial N=%d" % N, tag, len(r["cost_log"]), r["cost_log"][0], r["cost_log"][-1])
        out.update({"x_" + tag: r["x"][None], "u_" + tag: r["u"][None], "cost_log_" + tag: r["cost_log"][None]})
    np.savez_compressed(os.path.join(OUT, "tutorial_tassa_n%d.npz" % N), x0=p["x0"], u0=p["u0"], N=N, **out)


def golden_parking():
    """Parking between two cars (state projection onto obstacle sets): (a) the notebook configuration N=500, dt=0.03
    (known answer: first iterate 2564.0110889491493, Car/Iterative LQR with state constraints.ipynb cell 20 output),
    (b) five start states at N=200 whose paths cross the obstacles, (c) the reference's project_state closure on
    random trajectories."""
    out = {}
    p = P.parking_batch(1)
    r = S.run_parking(p, 0)
    out.update(nb_cost_log=r["cost_log"], nb_x=r["x"], nb_u=r["u"])
    print("parking notebook", r["cost_log"])
    p = P.parking_batch(5, N=200, dt=0.075)
    xs, us, logs = [], [], []
    for b in range(5):
        r = S.run_parking(p, b)
        xs.append(r["x"]); us.append(r["u"]); logs.append(r["cost_log"])
        print("parking", b, len(r["cost_log"]), r["cost_log"][-1])
    out.update(x=np.stack(xs), u=np.stack(us), cost_log=_pad(logs), x0=p["x0"])
    rng = np.random.default_rng(77)
    proj = S.parking_project_state(p)
    pts = np.zeros((6, 200, 4))
    pts[:, :, :2] = rng.uniform(-10.0, 0.0, (6, 200, 2))
    pts[:, :, 2:] = rng.normal(0, 1, (6, 200, 2))
    out.update(proj_in=pts, proj_out=np.stack([proj(q.flatten()).reshape(200, 4) for q in pts]))
    np.savez_compressed(os.path.join(OUT, "parking_obstacles.npz"), **out)


def golden_isls_admm(B=3):
    """Robust iSLS-ADMM of the unmodified reference (isls.py:503-712) on the 3-DoF arm with chance-constrained control
    bounds (3DoF robot/State bounds and robust control bounds.ipynb cells 12-26): problem 0 is the notebook's q0."""
    p = P.arm_robust_batch(B)
    model = M.make_model("arm3", dt=p["dt"])
    xs, us, logs, dus, phis = [], [], [], [], []
    for b in range(B):
        r = S.run_isls_admm(model, p, b)
        xs.append(r["x"]); us.append(r["u"]); logs.append(r["cost_log"]); dus.append(r["du"]); phis.append(r["phi_u"])
        print("isls_admm", b, len(r["cost_log"]), r["cost_log"][-1], np.abs(r["u"]).max())
    # controller + Monte-Carlo evaluation of the robust solution of problem 0 (notebook cells 21, 26): the legacy
    # `sls.controller(PHI_U, du)` of the notebook is SLS.controller (sls.py:235-242) on the iSLS operators C, D
    p0 = P.arm_robust_batch(1)
    s = S.make_isls(model, p0["N"], p0["zs"], _Qs(p0), p0["seq"], p0["u_std"])
    S.init_nominal(s, p0["x0"][0], p0["u0"])
    rb = p0["robust"]
    from isls.projections import project_set_convex, project_soc_unit

    def project_u(u, u_nom):
        y_ = u.copy()
        y_[:, 0] += u_nom.flatten()
        y_ = project_set_convex(y_, rb["As"], rb["bs"], projections=[project_soc_unit] * 2, rho=rb["inner_rho"],
                                max_iter=rb["inner_max_iter"], threshold=rb["inner_threshold"], verbose=0)
        y_[:, 0] -= u_nom.flatten()
        return y_
    with S.quiet():
        du, phi_u = s.isls_admm(3, model.get_AB, max_line_search=p0["L"], k_max=p0["I_o"], project_u=project_u,
                                rho_u=rb["rho_u"], max_admm_iter=p0["I_a"], threshold=p0["tol"], verbose=0, log=True)
        Nm, Nn = p0["N"] * p0["m"], p0["N"] * p0["n"]
        PHI_U = np.zeros((Nm, Nn))
        PHI_U[:, :3] = phi_u
        s.Sw, s.Su = s.C, s.D
        pkg, _ = S.load()
        K, k = pkg.SLS.controller(s, PHI_U, du)
        rng = np.random.default_rng(21)
        x0s = np.tile(s.x_nom[0:1], (64, 1))
        x0s[:, :3] = rng.normal(loc=s.x_nom[0, :3], scale=np.sqrt(0.1), size=(64, 3))       # cell 21
        # one sample per call: HEAD returns only the first trajectory for a 2-D x0 (isls_base.py:40-43)
        res = [s.get_trajectory_sls(x0, K, k, noise_scale=0) for x0 in x0s]
        xl, ul = np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res])
    print("controller: max|K - PHI_U|", np.abs(K - PHI_U).max(), " max|k - du|", np.abs(k - du).max(),
          " success %", 100 * np.mean(np.all(np.abs(ul) <= 6.0 + 1e-3, axis=(1, 2))))
    ctrl = dict(mc_x0=x0s, mc_u=ul, mc_x=xl, ctrl_K_minus_PHI=np.abs(K - PHI_U).max(), ctrl_k_minus_du=np.abs(k - du).max())
    # the same problems without any projection (notebook cell 23: isls_admm(q_dim, get_AB, max_line_search=10, ...))
    ulogs, uphis = [], []
    for b in range(B):
        s = S.make_isls(model, p["N"], p["zs"], _Qs(p), p["seq"], p["u_std"])
        S.init_nominal(s, p["x0"][b], p["u0"])
        with S.quiet():
            _, phi = s.isls_admm(3, model.get_AB, max_line_search=10, k_max=100, max_admm_iter=10, threshold=1e-4,
                                 verbose=0, log=True)
        ulogs.append(np.array(s.cost_log)); uphis.append(phi)
        print("isls_admm unconstrained", b, len(s.cost_log), s.cost_log[-1])
    np.savez_compressed(os.path.join(OUT, "arm_isls_admm.npz"), x0=p["x0"], x=np.stack(xs), u=np.stack(us),
                        cost_log=_pad(logs), du=np.stack(dus), phi_u=np.stack(phis), unc_cost_log=_pad(ulogs),
                        unc_phi_u=np.stack(uphis), **ctrl)


def golden_isls_admm_x(B=3):
    """Robust iSLS-ADMM of the unmodified reference with a STATE-side projection (isls.py:631-638: project_x on
    [d_x | Phi_x(:, :3)]): chance-constrained bounds on two joint velocities of the arm beside the control bounds
    (problems of configs.arm_robust_x_batch, closure in ref_shim.run_isls_admm), and once with project_x alone."""
    p = P.arm_robust_x_batch(B)
    model = M.make_model("arm3", dt=p["dt"])
    xs, us, logs, dus, phis = [], [], [], [], []
    for b in range(B):
        r = S.run_isls_admm(model, p, b)
        xs.append(r["x"]); us.append(r["u"]); logs.append(r["cost_log"]); dus.append(r["du"]); phis.append(r["phi_u"])
        print("isls_admm_x", b, len(r["cost_log"]), r["cost_log"][-1], "max|qd|", np.abs(r["x"][:, 3:6]).max(0))
    p1 = P.arm_robust_x_batch(1, project_u=False)
    r1 = S.run_isls_admm(model, p1, 0)
    print("isls_admm_x (state side only)", len(r1["cost_log"]), r1["cost_log"][-1], "max|qd|", np.abs(r1["x"][:, 3:6]).max(0),
          "max|u|", np.abs(r1["u"]).max())
    np.savez_compressed(os.path.join(OUT, "arm_isls_admm_x.npz"), x0=p["x0"], x=np.stack(xs), u=np.stack(us),
                        cost_log=_pad(logs), du=np.stack(dus), phi_u=np.stack(phis), xonly_x=r1["x"], xonly_u=r1["u"],
                        xonly_cost_log=r1["cost_log"], xonly_du=r1["du"], xonly_phi_u=r1["phi_u"])


def golden_di_obstacles():
    """LQT-ADMM with the spherical-obstacle state projection (set-convex + Dykstra) of the unmodified reference:
    the notebook problem with ADMM_LQT_DP (500 iterations: printed cost 2.701e-01) and ADMM_LQT_Batch, plus three more
    start states; residual logs kept in full (the non-convex iteration does not converge and amplifies rounding, so
    parity is checked on the first iterations)."""
    out = {}
    p = P.di_obstacle_batch(4, max_iter=500, tol=1e-4)
    res = [S.run_di_obstacles(p, b, "dp") for b in range(4)]
    out.update(x0=p["x0"], dp_x=np.stack([r["x"] for r in res]), dp_u=np.stack([r["u"] for r in res]),
               dp_logs=np.stack([r["logs"] for r in res]), dp_cost=np.array([r["cost"] for r in res]))
    print("di obstacles dp costs", out["dp_cost"])
    p = P.di_obstacle_batch(4, max_iter=200, tol=1e-3)
    res = [S.run_di_obstacles(p, b, "batch") for b in range(4)]
    out.update(batch_logs=np.stack([r["logs"] for r in res]), batch_cost=np.array([r["cost"] for r in res]))
    # the projection closure alone on random position clouds around the obstacles
    rng = np.random.default_rng(5)
    proj = S.di_obstacle_project_state(p)
    pts = np.zeros((5, 100, 4))
    pts[:, :, :2] = rng.uniform(0.1, 0.9, (5, 100, 2))
    pts[:, :, 2:] = rng.normal(0, 1, (5, 100, 2))
    out.update(proj_in=pts, proj_out=np.stack([proj(q.flatten()).reshape(100, 4) for q in pts]))
    np.savez_compressed(os.path.join(OUT, "di_obstacles.npz"), **out)


def golden_replan():
    """SLS.initialize_replanning_procedure / replan_feedforward (isls/sls.py:244-248) of the unmodified reference on the
    C4-style double integrator (n=4, m=2, N=50): controller from solve_sls, then three new targets."""
    pkg, _ = S.load()
    from isls.utils import get_double_integrator_AB
    n, m, N = 4, 2, 50
    A, B = get_double_integrator_AB(2, nb_deriv=2, dt=1.0 / N)
    zs = np.stack([np.zeros(n), np.array([0.8, 0.7, 0.0, 0.0])])
    Qs = np.stack([np.zeros((n, n)), np.eye(n) * 1e6])
    seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
    with S.quiet():
        s = pkg.SLS(n, m, N)
        s.AB = [A, B]
        s.set_quadratic_cost(zs, Qs, seq, 1e-2)
        PHI_U, du = s.solve_sls()
        K, k = s.controller(PHI_U, du)
        s.initialize_replanning_procedure(K)
        rng = np.random.default_rng(3)
        tg = rng.uniform(0.5, 1.0, (3, 2))
        xds, ks = [], []
        for t in tg:
            xd = s.xd.copy()
            xd[-n:-n + 2] = t
            xds.append(xd); ks.append(s.replan_feedforward(k, xd))
    np.savez_compressed(os.path.join(OUT, "sls_replan.npz"), zs=zs, K=K, k=k, xd_new=np.stack(xds), k_new=np.stack(ks),
                        xd_old=np.array(s.xd))


def golden_sls_state_bounds():
    """SLS.ADMM_SLS with project_u AND project_x of the unmodified reference on the notebook problem
    (Double integrator/LQR and SLS with state bounds.ipynb cells 16-17; printed: "primal: 1.37e-04 dual: 1.80e-01")."""
    pkg, _ = S.load()
    from isls.utils import get_double_integrator_AB
    from isls.projections import project_set_convex, project_soc_unit
    p = P.sls_state_bounds_problem()
    N, n = p["N"], p["n"]
    A, B = get_double_integrator_AB(1, nb_deriv=2, dt=p["dt"])
    kw = dict(rho=p["inner_rho"], max_iter=p["inner_max_iter"], threshold=p["inner_threshold"])
    project_u = lambda y: project_set_convex(y, p["As"], p["bs_u"], projections=[project_soc_unit] * 2, verbose=False, **kw)

    def project_x(x):
        x_ = x.copy()
        for row, As_x, bs_x in p["x_rows"]:
            x_[row:row + 1] = project_set_convex(x_[row:row + 1], As_x, bs_x, projections=[project_soc_unit] * 2, **kw)
        return x_
    rho_x = np.zeros((N, n, n))
    rho_x[-1, 0, 0] = rho_x[-1, 1, 1] = 1e3
    with S.quiet():
        s = pkg.SLS(n, 1, N)
        s.AB = [A, B]
        s.set_quadratic_cost(p["zs"], np.zeros((2, n, n)), p["seq"], p["u_std"])
        du, PHI, log = s.ADMM_SLS(project_u=project_u, project_x=project_x, max_iter=p["max_iter"], rho_x=rho_x,
                                  rho_u=p["rho_u"], alpha=1.0, tol=p["tol"], verbose=0, log=True)
    print("sls state bounds", len(log), log[-1])
    np.savez_compressed(os.path.join(OUT, "sls_state_bounds.npz"), du=du, phi_u=PHI, logs=np.array(log))


def golden_arm_lq_step(nb=3):
    """One outer x one ADMM x one candidate (alpha = 1) iteration of the unmodified reference on arm problems:
    u_head = u^ + du* with du* from HEAD's explicit dense inverse (isls.py:462-465), plus the condition number of the
    dense normal matrix.  Arbitrated against a 40-digit solve in tests/test_gpu_baseline_sizes.py."""
    from oracle import restated as R
    p = P.arm_batch(nb, I_o=1, I_a=1, L=1)
    model = M.make_model(p["model"], dt=p["dt"])
    us, conds = [], []
    for b in range(nb):
        s = S.make_isls(model, p["N"], p["zs"], _Qs(p), p["seq"], p["u_std"])
        S.init_nominal(s, p["x0"][b], p["u0"])
        rho_x = np.stack([np.diag(r) for r in p["rho_x"]])
        r = S.run_ilqr_admm(s, model, project_x=_clip(p["lo_x"], p["hi_x"]), project_u=_clip(p["lo_u"], p["hi_u"]),
                            rho_x=rho_x, rho_u=float(p["rho_u"][0, 0]), max_iter=1, max_admm_iter=1,
                            max_line_search_iter=1, tol=0.0)
        us.append(r["u"])
        # dense normal matrix Su'(Q + Qr)Su + R + Rr of this linearisation (what HEAD inverts)
        Qd = np.diag((p["Qdiag"][p["seq"]] + p["rho_x"]).reshape(-1))
        Rd = np.diag((p["u_std"] + p["rho_u"]).reshape(-1))
        conds.append(np.linalg.cond(s.Su.T @ Qd @ s.Su + Rd))
        print("arm_lq_step", b, "cond %.2e" % conds[-1])
    np.savez_compressed(os.path.join(OUT, "arm_lq_step.npz"), x0=p["x0"], u_head=np.stack(us), cond=np.array(conds))


def golden_controller_tv():
    """iSLS.controller of the unmodified reference (SLS.controller, sls.py:235-242, on the C, D that iSLSBase.AB builds,
    isls_base.py:138-158) for a GENERAL causal Phi_u - every block column populated, not only the first one that
    isls_admm returns - on the car (N = 20) and the arm (N = 12) around a rolled-out nominal trajectory."""
    out = {}
    pkg, _ = S.load()
    for name, p, N in (("car", P.car_batch(1), 20), ("arm", P.arm_batch(1), 12)):
        model = M.make_model(p["model"], dt=p["dt"])
        n, m = model.n, model.m
        rng = np.random.default_rng(77 + N)
        zs = p["zs"] if p["zs"].ndim == 2 else p["zs"][0]
        seq = np.zeros(N, dtype=np.int64)
        seq[-1] = p["seq"][-1]
        s = S.make_isls(model, N, zs, _Qs(p), seq, p["u_std"])
        u0 = rng.normal(scale=0.3, size=(N, m))
        S.init_nominal(s, p["x0"][0], u0)
        with S.quiet():
            A, Bm = model.get_AB(s.x_nom, s.u_nom)
            s.AB = A, Bm
            PHI_U = np.zeros((N * m, N * n))
            for t in range(N):
                PHI_U[t * m:(t + 1) * m, :(t + 1) * n] = rng.normal(scale=0.2, size=(m, (t + 1) * n))
            du = rng.normal(size=N * m)
            s.Sw, s.Su = s.C, s.D
            K, k = pkg.SLS.controller(s, PHI_U, du)
        print("controller_tv", name, "max|K|", np.abs(K).max(), "cond(PHI_X)", np.linalg.cond(s.C + s.D @ PHI_U))
        out.update({name + "_x": s.x_nom.copy(), name + "_u": s.u_nom.copy(), name + "_A": np.asarray(A),
                    name + "_B": np.asarray(Bm), name + "_PHI_U": PHI_U, name + "_du": du, name + "_K": K, name + "_k": k,
                    name + "_dt": np.array(p["dt"])})
    np.savez_compressed(os.path.join(OUT, "controller_tv.npz"), **out)


def golden_solve_dp_ff():
    """SLS.solve_dp(return_Qs=True) and SLS.solve_dp_ff of the unmodified reference (isls/sls.py:85-202) on the C1 double
    integrator: unregularised, and the regularised (ADMM) form with diagonal Qr, Rr and random xr, ur."""
    pkg, _ = S.load()
    p = P.di_batch(1)
    N, n, m = p["N"], p["n"], p["m"]
    model = M.make_model("double_integrator", nb_dim=m, dt=p["dt"])
    rng = np.random.default_rng(77)
    with S.quiet():
        s = pkg.SLS(n, m, N)
        s.AB = [model.A, model.B]
        s.set_quadratic_cost(p["zs"], _Qs(p), p["seq"], p["u_std"])
        K, k, Quu, Qui, Qux = s.solve_dp(return_Qs=True)
        k_ff = s.solve_dp_ff(K, Quu, Qux, Qui)
        qr, rr = rng.uniform(0.1, 2.0, (N, n)), rng.uniform(0.01, 1.0, (N, m))
        xr, ur = rng.normal(0, 0.5, N * n), rng.normal(0, 1.0, N * m)
        Qr = np.stack([np.diag(q) for q in qr])
        Rr = [np.diag(r) for r in rr]
        Kr, kr, Quur, Quir, Quxr = s.solve_dp(Qr=Qr, Rr=Rr, ur=ur, xr=xr, return_Qs=True)
        xr2, ur2 = rng.normal(0, 0.5, N * n), rng.normal(0, 1.0, N * m)
        kr_ff = s.solve_dp_ff(Kr, Quur, Quxr, Quir, Qr=Qr, Rr=Rr, ur=ur2, xr=xr2)
    np.savez_compressed(os.path.join(OUT, "di_solve_dp_ff.npz"), K=K, k=k, Quu=Quu, Quu_inv=Qui, Qux=Qux, k_ff=k_ff,
                        qr=qr, rr=rr, xr=xr, ur=ur, Kr=Kr, kr=kr, Quur=Quur, Quu_invr=Quir, Quxr=Quxr, xr2=xr2, ur2=ur2,
                        kr_ff=kr_ff)
    print("di_solve_dp_ff: |k_ff - k| %.2e" % np.abs(k_ff - k).max())


def golden_admm_toy():
    """The reference's generic ADMM driver (isls/admm.py:6-106) on a toy separable problem with a closed-form argmin:
    min |x - a|^2 + |u - b|^2  s.t. box bounds, f_argmin(reg) = (a + rho reg) / (1 + rho); alpha = 1 and 1.5."""
    pkg, _ = S.load()
    from isls.admm import ADMM
    rng = np.random.default_rng(5)
    a, b = rng.normal(0, 1.0, 12), rng.normal(0, 2.0, 8)
    rho = 0.7
    out = dict(a=a, b=b, rho=rho)
    for tag, alpha in (("a10", 1.0), ("a15", 1.5)):
        def f_argmin(reg_x, reg_u):
            return (a + rho * reg_x) / (1 + rho), (b + rho * reg_u) / (1 + rho)
        with S.quiet():
            r = ADMM(12, 8, f_argmin, project_x=lambda z: np.clip(z, -0.5, 0.5), project_u=lambda z: np.clip(z, -1.0, 1.5),
                     max_iter=200, alpha=alpha, tol=1e-6, return_lmb=True, log=True)
        x, u, lx, lu, zx, zu, logs = r
        out.update({tag + "_x": x, tag + "_u": u, tag + "_lx": lx, tag + "_lu": lu, tag + "_zx": zx, tag + "_zu": zu,
                    tag + "_logs": np.array(logs)})
        print("admm_toy", tag, len(logs))
    np.savez_compressed(os.path.join(OUT, "admm_toy.npz"), **out)


def golden_projections_ex():
    """project_multilinear, project_affine, project_soc, project_block_lower_triangular of the unmodified reference
    (isls/projections.py:46-68, 163-232, 277-286) on seeded rows."""
    S.load()
    from isls.projections import (project_affine, project_block_lower_triangular, project_multilinear, project_soc)
    rng = np.random.default_rng(21)
    x = rng.normal(0, 2.0, (64, 5))
    A = rng.normal(0, 1.0, (3, 5))
    l, u = np.array([-0.5, -1.0, 0.2]), np.array([0.5, 0.3, 0.9])
    ml = np.stack([project_multilinear(x[i], A, l, u) for i in range(64)])
    a = rng.normal(0, 1.0, 5)
    aff = np.stack([project_affine(x[i].copy(), a, 0.7, -0.4, 0.6) for i in range(64)])
    z0 = rng.normal(0, 1.5, (40, 3))
    As = rng.normal(0, 1.0, (4, 3))
    bs = np.array([0.1, -0.2, 0.0, 0.8])
    with S.quiet():
        soc = project_soc(z0.copy(), As, bs, rho=2.0, max_iter=100, tol=1e-6)
        soc1 = project_soc(z0[3].copy(), As, bs, rho=2.0, max_iter=100, tol=1e-6)
    N, xd, ud = 6, 4, 2
    Z = rng.normal(0, 1.0, (N * ud, N * xd))
    blt = project_block_lower_triangular(Z.copy(), xd, ud, N)
    np.savez_compressed(os.path.join(OUT, "projections_ex.npz"), x=x, A=A, l=l, u=u, ml=ml, a=a, aff=aff, z0=z0, As=As,
                        bs=bs, soc=soc, soc1=soc1, Z=Z, blt=blt)
    print("projections_ex: soc moved rows by up to %.3f" % np.abs(soc - z0).max())


def golden_lqt_lti():
    """SLS.ADMM_LQT_DP / ADMM_LQT_Batch / solve(method='dp') of the unmodified reference (isls/sls.py:40-60, 250-317) with a
    GENERAL constant pair (A, B) - a damped, coupled 2-D oscillator, not a double integrator - state and control bounds."""
    pkg, _ = S.load()
    N, n, m, dt = 40, 4, 2, 0.05
    rng = np.random.default_rng(31)
    A = np.eye(n) + dt * np.array([[0, 0, 1, 0], [0, 0, 0, 1], [-1.5, 0.4, -0.3, 0.1], [0.3, -2.0, 0.05, -0.2]])
    Bm = dt * np.array([[0.0, 0.0], [0.0, 0.0], [1.0, 0.2], [-0.1, 0.8]]) + 0.5 * dt * dt * rng.normal(0, 1, (n, m))
    zs = np.array([np.zeros(n), [0.6, -0.4, 0.0, 0.0]])
    Qs = np.stack([np.zeros((n, n)), np.diag([1e3, 1e3, 1e1, 1e1])])
    seq = np.zeros(N, dtype=np.int32); seq[-1] = 1
    x0s = np.array([[0.05, 0.0, 0.0, 0.1], [0.2, -0.1, 0.2, 0.0], [-0.3, 0.2, 0.0, -0.2]])
    lo_u, hi_u = np.full(N * m, -1.0), np.full(N * m, 1.0)
    lo_x, hi_x = np.full((N, n), -np.inf), np.full((N, n), np.inf)
    lo_x[:, 2:], hi_x[:, 2:] = -0.3, 0.3
    rho_x = np.stack([np.diag([0.0, 0.0, 1.0, 1.0])] * N)
    out = dict(A=A, B=Bm, zs=zs, Qdiag=np.stack([np.diag(q) for q in Qs]), seq=seq, x0=x0s, lo_x=lo_x, hi_x=hi_x)
    xs, us, Ks, its, xb, ub, itb, xu, uu = [], [], [], [], [], [], [], [], []
    for b in range(len(x0s)):
        with S.quiet():
            s = pkg.SLS(n, m, N)
            s.AB = [A, Bm]
            s.set_quadratic_cost(zs, Qs, seq, 1e-2)
            Kq, kq = s.solve(method="dp")
            xq, uq = s.get_trajectory_dp(x0s[b], Kq, kq)
            r = s.ADMM_LQT_DP(x0s[b], project_x=_clip(lo_x, hi_x), project_u=lambda z: np.clip(z, lo_u, hi_u), max_iter=600,
                              rho_x=rho_x, rho_u=1e-1, tol=1e-4, log=True)
            rb = s.ADMM_LQT_Batch(x0s[b], project_x=_clip(lo_x, hi_x), project_u=lambda z: np.clip(z, lo_u, hi_u),
                                  max_iter=600, rho_x=rho_x, rho_u=1e-1, tol=1e-4, log=True)
        xs.append(r[0].reshape(N, n)); us.append(r[1].reshape(N, m)); Ks.append(r[2]); its.append(len(r[-1]))
        xb.append(rb[0].reshape(N, n)); ub.append(rb[1].reshape(N, m)); itb.append(len(rb[-1]))
        xu.append(np.asarray(xq).reshape(N, n)); uu.append(np.asarray(uq).reshape(N, m))
        print("lqt_lti", b, its[-1], itb[-1], "max|u| %.3f max|v| %.3f" % (np.abs(us[-1]).max(), np.abs(xs[-1][:, 2:]).max()))
    out.update(x=np.stack(xs), u=np.stack(us), K=np.stack(Ks), iters=np.array(its), x_batch=np.stack(xb),
               u_batch=np.stack(ub), iters_batch=np.array(itb), x_unc=np.stack(xu), u_unc=np.stack(uu))
    np.savez_compressed(os.path.join(OUT, "lqt_lti.npz"), **out)


def golden_set_convex():
    """project_set_convex of the unmodified reference (isls/projections.py:289-374) over three sets of different kinds: a
    box on x itself, a quadratic shell of a 2-D linear image, a second-order cone of a 4-D affine image."""
    S.load()
    from isls.projections import project_quadratic_batch, project_set_convex, project_soc_unit
    rng = np.random.default_rng(41)
    x0 = rng.normal(0, 1.5, (50, 3))
    A0, b0 = np.eye(3), np.zeros(3)
    lo, hi = np.array([-0.8, -1.0, -0.5]), np.array([0.9, 0.7, 1.2])
    A1, b1 = rng.normal(0, 1.0, (2, 3)), np.array([0.1, -0.2])
    c1 = np.array([0.3, -0.1])
    A2, b2 = rng.normal(0, 0.7, (4, 3)), np.array([0.0, 0.1, -0.1, 1.5])
    projs = [lambda y: np.clip(y, lo, hi), lambda y: project_quadratic_batch(y - c1, 0.05, 0.8) + c1, project_soc_unit]
    with S.quiet():
        out = project_set_convex(x0.copy(), [A0, A1, A2], [b0, b1, b2], projs, rho=2.0, max_iter=150, threshold=1e-6)
    np.savez_compressed(os.path.join(OUT, "set_convex.npz"), x0=x0, A0=A0, b0=b0, lo=lo, hi=hi, A1=A1, b1=b1, c1=c1, A2=A2,
                        b2=b2, out=out)
    print("set_convex: moved rows by up to %.3f" % np.abs(out - x0).max())


if __name__ == "__main__":
    assert S.available(), "needs the reference tree"
    only = set(sys.argv[1:])                       # e.g. `make_golden.py tutorial` regenerates one fixture family

    def want(tag):
        return not only or tag in only
    if want("ilqr"):
        golden_ilqr_admm(P.car_batch(6), "car_ilqr_admm", 10.0)
        golden_ilqr_admm(P.car_batch(3, stress=True), "car_stress_ilqr_admm", 10.0)
        golden_ilqr_admm(P.arm_batch(3), "arm_ilqr_admm", 1e-3)
        golden_ilqr_dp(P.car_batch(4), "car_ilqr_dp", 30, 25)
        golden_ilqr_dp(P.arm_batch(2), "arm_ilqr_dp", 20, 25)
        golden_backward_pass("car_backward_pass", "car", 100, 11)
        golden_backward_pass("arm_backward_pass", "arm3", 100, 12)
        golden_lqt_admm_dp(P.di_batch(3), "di_lqt_admm_dp")
        golden_notebook_pins()
    if want("sls"):
        golden_sls()
    if want("mc"):
        golden_mc()
    if want("tutorial"):
        golden_tutorial()
    if want("tutorial_fullsize"):
        golden_tutorial_fullsize()
    if want("parking"):
        golden_parking()
    if want("isls_admm"):
        golden_isls_admm()
    if want("isls_admm_x"):
        golden_isls_admm_x()
    if want("sls_state"):
        golden_sls_state_bounds()
    if want("replan"):
        golden_replan()
    if want("di_obstacles"):
        golden_di_obstacles()
    if want("set_convex"):
        golden_set_convex()
    if want("lqt_lti"):
        golden_lqt_lti()
    if want("projections_ex"):
        golden_projections_ex()
    if want("admm_toy"):
        golden_admm_toy()
    if want("arm_lq_step"):
        golden_arm_lq_step()
    if want("solve_dp_ff"):
        golden_solve_dp_ff()
    if want("controller_tv"):
        golden_controller_tv()
    if want("lqt_batch"):
        golden_lqt_admm_batch(P.di_batch(3), "di_lqt_admm_batch")
