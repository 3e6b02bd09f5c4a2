"""TEST INFRASTRUCTURE (oracle) - drives the UNMODIFIED reference (`isls` at HEAD) from outside.

Only usable where the reference tree exists (this build container: /root/reference, or $ISLS_REFERENCE).
It is used solely by tests/golden/make_golden.py to produce the committed golden vectors and by the
container-only cross-check tests; nothing on the GPU box imports it (the reference does not travel).

The reference is imported read-only and patched from outside, never edited:
  S0  stub `matplotlib` package on sys.path (base.py:4, sls_base.py:4, utils.py:6-8 import it at module level)
  S1  obj.C = obj.Sw, obj.D = obj.Su            (isls_base.py:152-158 / isls.py:426-438 use C, D; base.py:18-19
                                                 defines Sw, Su)
  S2  ADMM(threshold=...) -> ADMM(tol=...)       (isls.py:480-486 passes `threshold`, admm.py:6-8 takes `tol`)
  S3  always log=True                            (isls.py:489-490 index the return tuple assuming the log entry)
  S4  quadratic via-point cost as cost_function  (isls_base.py:113-117 falls back to an undefined compute_cost)
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("ISLS_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "isls"))


_mod = {}


def load():
    """Import the reference package (once) and return (isls_module, isls.isls module)."""
    if _mod:
        return _mod["pkg"], _mod["isls"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    stubs = os.path.join(_HERE, "stubs")
    for p in (REFERENCE_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import isls as pkg                      # noqa: E402
        import isls.isls as isls_mod            # noqa: E402
    orig_admm = isls_mod.ADMM

    def admm_accepting_threshold(*a, threshold=None, **k):          # shim S2
        if threshold is not None:
            k["tol"] = threshold
        return orig_admm(*a, **k)

    isls_mod.ADMM = admm_accepting_threshold
    _mod.update(pkg=pkg, isls=isls_mod, orig_admm=orig_admm)
    return pkg, isls_mod


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def quadratic_cost(Q_dense, xd, R_dense):
    """sum (x-xd)' Q (x-xd) + u' R u over the flattened horizon, batched over a leading axis
    (same expression as sls_base.py:25-44)."""
    def cost(x, u):
        single = (x.ndim == 2)
        xf = x.reshape(1, -1) if single else x.reshape(x.shape[0], -1)
        uf = u.reshape(1, -1) if single else u.reshape(u.shape[0], -1)
        dx = xf - xd
        c = np.sum(dx * (Q_dense @ dx.T).T, axis=-1) + np.sum(uf * (R_dense @ uf.T).T, axis=-1)
        return c[0] if single else c
    return cost


def make_isls(model, N, zs, Qs, seq, u_std):
    """iSLS object with forward model, quadratic via-point cost and shims S1/S4 applied."""
    pkg, _ = load()
    with quiet():
        s = pkg.iSLS(model.n, model.m, N)
        s.C, s.D = s.Sw, s.Su                                        # shim S1
        s.forward_model = model.f
        s.set_quadratic_cost(zs, Qs, seq, u_std)
        Qd = s.Q.toarray()
        Rd = s.R.toarray()
        s.cost_function = quadratic_cost(Qd, s.xd, Rd)               # shim S4
    return s


def get_Cs_quadratic(s):
    """Analytic get_Cs for the via-point cost: c=[2Q_t(x-z_t); 2R u], C=blkdiag(2Q_t, 2R)
    (what isls.py:263-271 computes inline in the `Cts is None` branch)."""
    n, m, N = s.x_dim, s.u_dim, s.N

    def get_Cs(x, u):
        c = np.zeros((N, n + m))
        C = np.zeros((N, n + m, n + m))
        for t in range(N):
            Q = s.Qs[s.seq[t]]
            c[t, :n] = 2 * Q @ (x[t] - s.zs[s.seq[t]])
            c[t, n:] = 2 * s.Rt @ u[t]
            C[t, :n, :n] = 2 * Q
            C[t, n:, n:] = 2 * s.Rt
        return c, C
    return get_Cs


def init_nominal(s, x0, u0):
    with quiet():
        x_nom, u_nom = s.rollout_batch(np.asarray(x0)[None], np.asarray(u0)[None])
        s.reset()
        s.nominal_values = x_nom[0], u_nom[0]


def run_ilqr_admm(s, model, project_x=False, project_u=False, rho_x=None, rho_u=None, max_iter=20,
                  max_admm_iter=20, max_line_search_iter=20, alpha=1.0, tol=1e-3):
    """Shimmed HEAD `iSLS.ilqr_admm` (isls.py:379-501). Returns dict(x, u, cost_log, admm_log)."""
    with quiet():
        log = s.ilqr_admm(model.get_AB, project_x=project_x, project_u=project_u, rho_x=rho_x, rho_u=rho_u,
                          max_iter=max_iter, max_admm_iter=max_admm_iter,
                          max_line_search_iter=max_line_search_iter, alpha=alpha, tol=tol, log=True)  # S3
    return dict(x=s.x_nom.copy(), u=s.u_nom.copy(), cost_log=np.array(s.cost_log, dtype=np.float64),
                admm_log=np.array(log, dtype=np.float64))


def run_ilqr_dp(s, model, max_iter=100, max_line_search_iter=25, tol_fun=1e-5):
    """HEAD `iSLS.solve(method='dp')` (isls.py:54-132) with the analytic get_Cs."""
    with quiet():
        s.solve(model.get_AB, get_Cs_quadratic(s), method="dp", max_iter=max_iter,
                max_line_search_iter=max_line_search_iter, tol_fun=tol_fun)
    return dict(x=s.x_nom.copy(), u=s.u_nom.copy(), cost_log=np.array(s.cost_log, dtype=np.float64))


# ----------------------------------------------------------------------------- Tutorial problem (pseudo-Huber cost)
def tutorial_cost(p):
    """The Tutorial's cost closure (notebooks/Tutorial.ipynb cell 14) in plain numpy: lu + lf + lx per step, summed;
    NaN -> 1e6 for batched inputs.  p: oracle problem dict (configs.tassa_batch)."""
    N = p["N"]
    cu = np.asarray(p["Rdiag"])
    cf_ = np.asarray(p["Qdiag_b"])[p["seq"]]                       # [N, n]: zero rows except the last
    pf = np.asarray(p["Hp_b"])[1]
    cx_ = np.asarray(p["Qdiag"])[p["seq"]][:, :2]
    px = np.asarray(p["Hp"])[0][:2][None]

    def pseudo_huber(x, pp):
        return np.sqrt(x ** 2 + pp ** 2) - pp

    def cost_vec(x, u):
        lu = np.sum(cu * (u ** 2), axis=-1)
        lf = cf_ @ pseudo_huber(x[-1], pf)
        lx = np.sum(cx_ * pseudo_huber(x[:, :2], px), axis=-1)
        return lf + lu + lx

    def cost(x, u):
        if x.ndim == 3:
            costs = np.zeros(x.shape[0])
            for i in range(x.shape[0]):
                costs[i] = np.sum(cost_vec(x[i], u[i]), -1)
            costs[np.isnan(costs)] = 1e6
        else:
            costs = np.sum(cost_vec(x, u), -1)
        return costs
    return cost


def tutorial_get_Cs(p):
    """Analytic replacement of the autograd get_Cs of Tutorial cell 16 (autograd is not installed here): gradient and
    (diagonal) Hessian of cost_vec; the cost has no x-u coupling (the notebook says so), so Cux = 0."""
    n, m, N = p["n"], p["m"], p["N"]
    terms = [(np.asarray(p["Qdiag"])[p["seq"]], np.asarray(p["Hp"])[p["seq"]]),
             (np.asarray(p["Qdiag_b"])[p["seq"]], np.asarray(p["Hp_b"])[p["seq"]])]
    cu = np.asarray(p["Rdiag"])

    def get_Cs(x, u):
        c = np.zeros((N, n + m))
        C = np.zeros((N, n + m, n + m))
        g = np.zeros((N, n))
        h = np.zeros((N, n))
        for W, P in terms:
            sq = np.sqrt(x * x + P * P)
            g += W * x / sq
            h += W * (P * P) / (sq * sq * sq)
        c[:, :n] = g
        c[:, n:] = 2.0 * cu * u
        i = np.arange(n)
        C[:, i, i] = h
        j = n + np.arange(m)
        C[:, j, j] = 2.0 * cu
        return c, C
    return get_Cs


def make_isls_tutorial(model, p):
    """iSLS object set up the way Tutorial cells 17-18 do (forward_model + cost_function callables)."""
    pkg, _ = load()
    with quiet():
        s = pkg.iSLS(model.n, model.m, p["N"])
        s.C, s.D = s.Sw, s.Su                                        # shim S1
        s.forward_model = model.f
        s.cost_function = tutorial_cost(p)
    return s


def run_tutorial_dp(s, model, p, max_iter=100, max_line_search_iter=40):
    with quiet():
        s.solve(model.get_AB, tutorial_get_Cs(p), max_iter=max_iter, max_line_search_iter=max_line_search_iter,
                method="dp", verbose=False)                           # Tutorial cell 20
    return dict(x=s.x_nom.copy(), u=s.u_nom.copy(), cost_log=np.array(s.cost_log, dtype=np.float64))


def run_tutorial_admm(s, model, p):
    lo, hi = p["lo_u"].flatten(), p["hi_u"].flatten()
    with quiet():
        log = s.ilqr_admm(get_AB=model.get_AB, get_Cs=tutorial_get_Cs(p), project_u=lambda z: np.clip(z, lo, hi),
                          max_iter=p["I_o"], max_admm_iter=p["I_a"], max_line_search_iter=p["L"],
                          rho_u=np.diag(p["rho_u"][0]), tol=p["tol"], verbose=False, log=True)   # Tutorial cell 27
    return dict(x=s.x_nom.copy(), u=s.u_nom.copy(), cost_log=np.array(s.cost_log, dtype=np.float64),
                admm_log=np.array(log, dtype=np.float64))


# ------------------------------------------------------- parking between two cars (state constraints, obstacle sets)
def parking_project_state(p):
    """The notebook's project_state closure (Car/Iterative LQR with state constraints.ipynb cell 18), built from the
    reference's own project_set_convex / project_square_batch: the position must stay outside two rotated rectangles
    (infinity-norm shells in the frame W_i around the parked cars)."""
    load()
    from isls.projections import project_set_convex, project_square_batch
    ob = p["obstacles"]
    N, d = p["N"], p["n"]
    Ws, Ws_inv, xs_ = ob["W"], ob["W_inv"], ob["centers"]
    lower_sq, upper_sq = ob["lower"], ob["upper"]

    def make_function(i):
        def f(y):
            y_ = y.reshape(N, d).copy()
            z = y_[:, :2] - xs_[i][None]
            z_projected = project_square_batch(z @ Ws[i].T, lower_sq[i], upper_sq)
            z_projected = z_projected @ Ws_inv[i].T
            y_[:, :2] = z_projected + xs_[i][None]
            return y_
        return f

    projections = [make_function(i) for i in range(len(xs_))]
    As = [np.eye(d)] * len(xs_)
    bs = [np.zeros(d)] * len(xs_)

    def project_state(x):
        x_ = x.reshape(N, d).copy()
        return project_set_convex(x_, As, bs, projections, rho=ob["rho"], max_iter=ob["max_iter"], verbose=0,
                                  threshold=ob["threshold"]).flatten()
    return project_state


def run_parking(p, b=0):
    """Shimmed HEAD ilqr_admm on the parking problem (cell 20: state projection only)."""
    from . import models as M
    model = M.make_model("car", dt=p["dt"])
    s = make_isls(model, p["N"], p["zs"], np.stack([np.diag(q) for q in p["Qdiag"]]), p["seq"], p["u_std"])
    init_nominal(s, p["x0"][b], p["u0"])
    rho_x = np.stack([np.diag(r) for r in p["rho_x"]])
    return run_ilqr_admm(s, model, project_x=parking_project_state(p), rho_x=rho_x, max_iter=p["I_o"],
                         max_admm_iter=p["I_a"], max_line_search_iter=p["L"], tol=p["tol"])


# ------------------------------------------------------------ robust iSLS-ADMM (3-DoF arm, chance-constrained controls)
def soc_chance_cones(dim, var_x0, prob, lower_u, upper_u):
    """The A_, b_ of `3DoF robot/State bounds and robust control bounds.ipynb` cell 24 (b_ as in `Double
    integrator/LQR and SLS with control bounds.ipynb` cell 15, where the same construction is spelled out):
    rows z = [d_u + u_nom | Phi_u(:, :dim)] must satisfy  psi^-1 ||sqrt(sigma) z|| <= upper - z mu  and the mirrored
    lower bound."""
    from scipy.stats import norm
    mu = np.zeros(1 + dim)
    mu[0] = 1.0
    sigma = np.zeros(1 + dim)
    sigma[1:] = var_x0
    psi_inv = norm.ppf(prob)
    Au = np.diag(np.sqrt(sigma))
    As = [np.concatenate([Au, (-mu / psi_inv)[None]], axis=0), np.concatenate([Au, (mu / psi_inv)[None]], axis=0)]
    bs = [np.append(np.zeros(1 + dim), upper_u / psi_inv), np.append(np.zeros(1 + dim), -lower_u / psi_inv)]
    return As, bs


def run_isls_admm(model, p, b=0):
    """Shimmed HEAD `iSLS.isls_admm` (isls.py:503-712) on problem b of an oracle problem dict with p["robust"]."""
    load()
    from isls.projections import project_set_convex, project_soc_unit
    rb = p["robust"]
    s = make_isls(model, p["N"], p["zs"], np.stack([np.diag(q) for q in p["Qdiag"]]), p["seq"], p["u_std"])
    init_nominal(s, p["x0"][b], p["u0"])
    As, bs = rb["As"], rb["bs"]

    def project_u(u, u_nom):                                        # notebook cell 25
        u_nom_ = u_nom.flatten()
        y_ = u.copy()
        y_[:, 0] += u_nom_
        y_ = project_set_convex(y_, As, bs, projections=[project_soc_unit] * len(As), rho=rb["inner_rho"],
                                max_iter=rb["inner_max_iter"], threshold=rb["inner_threshold"], verbose=0)
        y_[:, 0] -= u_nom_
        return y_
    kw = dict(project_u=project_u, rho_u=rb["rho_u"])
    if rb.get("u_unprojected"):
        kw = {}
    rx = rb.get("x")
    if rx:
        N, n = p["N"], p["n"]

        def project_x(x, x_nom):                                    # the state-side twin of cell 25
            x_nom_ = x_nom.flatten()
            y_ = x.copy()
            y_[:, 0] += x_nom_
            for g, comp in enumerate(rx["comps"]):
                rows = np.arange(N) * n + comp
                y_[rows] = project_set_convex(y_[rows], As, list(rx["bs"][g]), projections=[project_soc_unit] * len(As),
                                              rho=rb["inner_rho"], max_iter=rb["inner_max_iter"],
                                              threshold=rb["inner_threshold"], verbose=0)
            y_[:, 0] -= x_nom_
            return y_
        kw.update(project_x=project_x, rho_x=np.diag(rx["rho_x"][0]))
    with quiet():
        du, phi_u = s.isls_admm(rb["dim"], model.get_AB, max_line_search=p["L"], k_max=p["I_o"],
                                max_admm_iter=p["I_a"], threshold=p["tol"], verbose=0, log=True, **kw)
    return dict(x=s.x_nom.copy(), u=s.u_nom.copy(), cost_log=np.array(s.cost_log, dtype=np.float64), du=du.copy(),
                phi_u=phi_u.copy())


# ------------------------------------------------------ double integrator, spherical obstacles (LQT-ADMM, state proj.)
def di_obstacle_project_state(p):
    """project_state of `Double integrator/LQR and SLS with spherical obstacle avoidance.ipynb` cell 12, built from the
    reference's own project_set_convex, project_quadratic and project_set_convex_dykstra."""
    load()
    from isls.projections import project_quadratic, project_set_convex, project_set_convex_dykstra
    ob = p["obstacles"]
    d, x_dim = p["n"], 2
    xs, lowers, upper = ob["centers"], ob["lower"], ob["upper"]
    K = len(xs)
    As = [np.eye(x_dim)] * K
    bs = [-xs[i] * 0 for i in range(K)]
    projections = [lambda x, lower=lower, i=i: project_quadratic(x - xs[i], lower, upper) + xs[i]
                   for i, lower in enumerate(lowers)]

    def project_state(x):
        x_ = x.reshape(-1, d).copy()
        x_[:, :x_dim] = project_set_convex(x_[:, :x_dim], As, bs, projections, max_iter=ob["max_iter"], verbose=0,
                                           threshold=ob["threshold"])
        x_[:, :x_dim] = project_set_convex_dykstra(x_[:, :x_dim], projections, max_iter=ob["dykstra_max_iter"],
                                                   verbose=0, tol=ob["dykstra_tol"])
        return x_.flatten()
    return project_state


def run_di_obstacles(p, b=0, form="dp"):
    from . import models as M
    pkg, _ = load()
    N, n, m = p["N"], p["n"], p["m"]
    model = M.make_model("double_integrator", nb_dim=m, dt=p["dt"])
    rho_x = np.stack([np.diag(r) for r in p["rho_x"]])
    with quiet():
        s = pkg.SLS(n, m, N)
        s.AB = [model.A, model.B]
        s.set_quadratic_cost(p["zs"], np.stack([np.diag(q) for q in p["Qdiag"]]), p["seq"], p["u_std"])
        fn = s.ADMM_LQT_DP if form == "dp" else s.ADMM_LQT_Batch
        r = fn(p["x0"][b], project_x=di_obstacle_project_state(p), max_iter=p["I_a"], rho_x=rho_x, alpha=1.0,
               tol=p["tol"], verbose=False, log=True)
        cost = s.compute_cost(r[0], r[1])
    return dict(x=r[0].reshape(N, n), u=r[1].reshape(N, m), logs=np.array(r[-1]), cost=float(cost))
