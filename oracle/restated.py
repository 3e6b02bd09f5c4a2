"""TEST INFRASTRUCTURE (oracle) - numpy restatement of the reference's iLQR / iLQR-ADMM / LQT-ADMM hot path.

This is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The CUDA path must not (and does not) route through it.

Parity status: PINNED.  tests/test_oracle_vs_reference.py (container only, needs /root/reference; skipped elsewhere) runs the
unmodified reference (oracle/ref_shim.py) side by side with this file, and tests/golden/*.npz hold
reference-generated vectors (made by tests/golden/make_golden.py) that this file is checked against
everywhere (including the GPU box, where the reference does not exist).

What is restated (reference file:line -> function here):
  isls/isls_base.py:10-11   alpha grid 10**linspace(0,-5,50)                     -> alphas()
  isls/sls_base.py:25-44    quadratic via-point cost (no 1/2)                    -> quad_cost()
  isls/isls.py:135-154      open-loop rollout for all candidates                 -> rollout_open()
  isls/isls.py:310-334      closed-loop rollout u=K(x-x^)+a k+u^                 -> rollout_closed()
  isls/isls.py:229-308      Riccati recursion K_t,k_t (dposv on Quu)             -> backward_pass()
  isls/sls.py:85-166        the same with ADMM regularisers + Quu/Qux logs       -> backward_pass(logs=True)
  isls/sls.py:168-202       feed-forward-only recursion (constant K)             -> ff_pass()
  isls/isls.py:336-374      iterate_once_dp: line search, NaN->1e5, accept test  -> ilqr_dp()
  isls/isls.py:54-132       iLQR outer loop and its stop rules                   -> ilqr_dp()
  isls/isls.py:379-501      ilqr_admm outer loop, f_argmin, warm start, stops    -> ilqr_admm()
  isls/admm.py:6-106        ADMM driver: z-projection, scaled dual, residuals    -> (inlined in ilqr_admm / lqt_admm_dp)
  isls/projections.py:7-11  project_bound = np.clip                              -> np.clip
  isls/sls.py:298-317       ADMM_LQT_DP (Riccati once, ff-pass per iteration)    -> lqt_admm_dp()
  isls/base.py:55-79        rho expansion (diagonal rho only)    -> problem dict rho_x / rho_u

The reference's `ilqr_admm` solves the inner LQ problem in dense batch least-squares form
(isls.py:436-465).  The recipe used here is the Riccati form of the same minimiser (SURVEY.md 8c'): one K-pass
per linearisation, one ff-pass + linear rollout per ADMM iteration giving the identical delta_u*, including the
batch form's treatment of the last control (u_{N-1} does not influence any state, so
delta_u_{N-1} = -Cuu^-1 cu_{N-1}).  It is checked against the unmodified reference in the container tests.

Everything is batched over a leading problem axis (the reference solves one problem; B=1 reproduces it).
Per-problem early exit is handled by compacting the still-active problems, so each problem sees exactly the
iteration sequence the reference would run for it alone.
"""
import numpy as np

from . import models as _models

# status bits (mirrored by include/isls_b200.h)
ST_CONVERGED_COST = 1      # |cost - prev_cost| < tol   (isls.py:125, isls.py:493)
ST_LINESEARCH_FAIL = 2     # forward pass failed         (isls.py:128)
ST_MAX_ITER = 4            # budget exhausted            (isls.py:131)
ST_OSCILLATING = 8         # mean-of-4 test              (isls.py:497)
ST_NON_PD = 16             # Quu not positive definite   (LinAlgError at isls.py:296)
ST_NAN_COST = 32           # NaN cost seen in line search (isls.py:362)

ADMM_CONVERGED = 1         # admm.py:72
ADMM_STALLED = 2           # admm.py:80
ADMM_MAXIT = 3             # admm.py:93


def alphas(L):
    """isls_base.py:10-11."""
    return (10.0 ** np.linspace(0.0, -5.0, 50))[:L]


# --------------------------------------------------------------------------------------------------- cost
def _zs_b(p, B):
    zs = np.asarray(p["zs"], dtype=np.float64)
    if zs.ndim == 2:
        zs = np.broadcast_to(zs, (B,) + zs.shape)
    return zs


def quad_cost(p, zs, x, u):
    """sum_t (x_t-z_t)'Q_t(x_t-z_t) + u_std * sum u^2  (sls_base.py:25-44 with block-diagonal Q from
    utils.py:101-115 and R = u_std I, base.py:86-89).  x[B,L,N,n], u[B,L,N,m], zs[B,k,n] -> [B,L]."""
    seq = p["seq"]
    Qd = p["Qdiag"][seq]                       # [N,n]
    xd = zs[:, seq]                            # [B,N,n]
    dx = x - xd[:, None]
    c = np.sum(dx * dx * Qd, axis=(-1, -2))
    c = c + p["u_std"] * np.sum(u * u, axis=(-1, -2))
    return c


def _R(p):
    """Diagonal of R: u_std * I (base.py:86-89) or the per-control weights of the Tutorial (cell 14: cu)."""
    if p.get("Rdiag") is not None:
        return np.asarray(p["Rdiag"], dtype=np.float64)
    return np.full(p["m"], float(p["u_std"]))


def _huber_terms(p):
    """[(W[N,n], P[N,n])]: the weighted pseudo-Huber terms per time step (running term a, final term b)."""
    seq = p["seq"]
    terms = [(np.asarray(p["Qdiag"], float)[seq], np.asarray(p["Hp"], float)[seq])]
    if p.get("Qdiag_b") is not None:
        terms.append((np.asarray(p["Qdiag_b"], float)[seq], np.asarray(p["Hp_b"], float)[seq]))
    return terms


def total_cost(p, zs, x, u):
    """cost_function of the problem: the quadratic via-point cost, or the Tutorial's pseudo-Huber cost
    (notebooks/Tutorial.ipynb cell 14: lu + lf + lx, NaN -> 1e6 for batched inputs)."""
    if p.get("cost", "quadratic") == "quadratic" and p.get("Rdiag") is None:
        return quad_cost(p, zs, x, u)
    e = x - zs[:, p["seq"]][:, None]
    if p.get("cost", "quadratic") == "quadratic":
        c = np.sum(e * e * np.asarray(p["Qdiag"], float)[p["seq"]], axis=(-1, -2))
    else:
        c = 0.0
        for W, P in _huber_terms(p):
            c = c + np.sum(W * (np.sqrt(e * e + P * P) - P), axis=(-1, -2))
        c = c + np.sum(_R(p) * u * u, axis=(-1, -2))
        return np.where(np.isnan(c), 1e6, c)
    return c + np.sum(_R(p) * u * u, axis=(-1, -2))


def state_grad_hess(p, zs_seq, x):
    """cts[:, :n] and diag(Cts[:, :n, :n]) of get_Cs at x[b,N,n] (Tutorial cell 16; isls.py:263-279): quadratic
    2Q(x-z), 2Q; pseudo-Huber w e / s, w p^2 / s^3, s = sqrt(e^2 + p^2)."""
    e = x - zs_seq
    if p.get("cost", "quadratic") == "quadratic":
        Qd = np.asarray(p["Qdiag"], float)[p["seq"]]
        return 2.0 * Qd * e, np.broadcast_to(2.0 * Qd, e.shape)
    g = np.zeros_like(e)
    h = np.zeros_like(e)
    for W, P in _huber_terms(p):
        sq = np.sqrt(e * e + P * P)
        g = g + W * e / sq
        h = h + W * (P * P) / (sq * sq * sq)
    return g, h


# ------------------------------------------------------------------------------------------------ rollouts
def rollout_open(model, x0, u):
    """isls.py:135-154.  x0[B,n], u[B,L,N,m] -> x[B,L,N,n]."""
    B, L, N, m = u.shape
    x = np.broadcast_to(x0[:, None, :], (B, L, x0.shape[-1])).copy()
    xs = np.empty((B, L, N, x0.shape[-1]))
    for t in range(N):
        xs[:, :, t] = x
        x = model.f(x, u[:, :, t])
    return xs


def rollout_closed(model, x_nom, u_nom, K, k_cand):
    """isls.py:310-334.  x_nom[B,N,n], u_nom[B,N,m], K[B,N,m,n], k_cand[B,L,N,m]."""
    B, L, N, m = k_cand.shape
    n = x_nom.shape[-1]
    x = np.broadcast_to(x_nom[:, None, 0, :], (B, L, n)).copy()
    xs = np.empty((B, L, N, n))
    us = np.empty((B, L, N, m))
    for t in range(N):
        dx = x - x_nom[:, None, t]
        u = np.einsum("bij,blj->bli", K[:, t], dx) + k_cand[:, :, t] + u_nom[:, None, t]
        us[:, :, t] = u
        xs[:, :, t] = x
        x = model.f(x, u)
    return xs, us


# ------------------------------------------------------------------------------------------------- Riccati
def _T(a):
    return np.swapaxes(a, -1, -2)


def backward_pass(A, Bm, cx, cu, Cxx, Cuu, Cux=None, logs=False, joseph=False):
    """Riccati recursion of isls.py:229-308 (general `Cts` branch), batched.

    A[B,N,n,n], Bm[B,N,n,m], cx[B,N,n], cu[B,N,m], Cxx[B,N,n,n], Cuu[B,N,m,m] (leading B may be 1 for the
    cost terms), Cux[B,N,m,n] or None.  Returns K[B,N,m,n], k[B,N,m], non_pd[B] (+ Quu, Quu_inv, Qux logs as in
    sls.py:117-120,159-162).  K[N-1] = k[N-1] = 0 (isls.py:245-246, 261).

    joseph=True evaluates the same V in the closed-loop form V = Cxx + K'Cuu K + (A + BK)'V(A + BK) (needs Cux = 0).
    The reference's four-term expression (isls.py:300) cancels catastrophically in FP64 when the control is cheap
    next to the accumulated state weights: on the arm with R = 1e-4 and Qr = 10 on the joint velocities (isls_admm with
    project_x alone) it reproduces the first controls of the exact minimiser - the reference's dense solve, a 60-digit
    Riccati recursion - to 1 digit only, the closed-loop form to 1e-11.  isls_admm, whose reference IS the dense solve
    (isls.py:562-579), uses it.
    """
    Bsz, N, n, m = Bm.shape
    K = np.zeros((Bsz, N, m, n))
    k = np.zeros((Bsz, N, m))
    non_pd = np.zeros(Bsz, dtype=bool)
    V = np.broadcast_to(Cxx[:, -1], (Bsz, n, n)).copy()           # isls.py:257
    v = np.broadcast_to(cx[:, -1], (Bsz, n)).copy()               # isls.py:258
    if logs:
        Quu_log = np.zeros((Bsz, N, m, m))
        Quu_inv_log = np.zeros((Bsz, N, m, m))
        Qux_log = np.zeros((Bsz, N, m, n))
    for t in range(N - 2, -1, -1):
        At, Bt = A[:, t], Bm[:, t]
        qx = cx[:, t] + np.einsum("bji,bj->bi", At, v)            # isls.py:285
        qu = cu[:, t] + np.einsum("bji,bj->bi", Bt, v)            # isls.py:286
        VA = V @ At
        Qxx = Cxx[:, t] + _T(At) @ VA                             # isls.py:288
        Qux = _T(Bt) @ VA                                         # isls.py:289
        if Cux is not None:
            Qux = Qux + Cux[:, t]
        Quu = Cuu[:, t] + _T(Bt) @ (V @ Bt)                       # isls.py:290
        # dposv (isls.py:296) raises for non-PD Quu: flag those problems and keep them finite
        ev = np.linalg.eigvalsh(0.5 * (Quu + _T(Quu)))
        bad = ~(ev.min(axis=-1) > 0.0)
        non_pd |= bad
        Quu_s = np.where(bad[:, None, None], np.eye(m), Quu)
        rhs = np.concatenate([Qux, qu[:, :, None]], axis=-1)
        sol = -np.linalg.solve(Quu_s, rhs)
        Kt, kt = sol[:, :, :-1], sol[:, :, -1]
        if joseph:
            Acl = At + Bt @ Kt
            Vn = _T(Acl) @ (V @ Acl) + _T(Kt) @ np.broadcast_to(Cuu[:, t], (Bsz, m, m)) @ Kt
            V = 0.5 * (Vn + _T(Vn)) + Cxx[:, t]
        else:
            V = Qxx + _T(Kt) @ Quu @ Kt + _T(Qux) @ Kt + _T(Kt) @ Qux                   # isls.py:300
        v = (qx + np.einsum("bji,bj->bi", Kt, qu) + np.einsum("bji,bj->bi", Kt, np.einsum("bij,bj->bi", Quu, kt))
             + np.einsum("bji,bj->bi", Qux, kt))                                         # isls.py:302
        K[:, t], k[:, t] = Kt, kt
        if logs:
            Quu_log[:, t], Qux_log[:, t] = Quu, Qux
            Quu_inv_log[:, t] = np.linalg.inv(Quu_s)
    if logs:
        return K, k, non_pd, Quu_log, Quu_inv_log, Qux_log
    return K, k, non_pd


def ff_pass(A, Bm, cx, cu, K, Quu, Quu_inv, Qux):
    """Feed-forward-only recursion of sls.py:168-202 (time-varying A,B).  Returns k[B,N,m] with k[N-1]=0."""
    Bsz, N, n, m = Bm.shape
    k = np.zeros((Bsz, N, m))
    v = cx[:, -1].copy()
    for t in range(N - 2, -1, -1):
        qx = cx[:, t] + np.einsum("bji,bj->bi", A[:, t], v)                              # sls.py:196
        qu = cu[:, t] + np.einsum("bji,bj->bi", Bm[:, t], v)                             # sls.py:197
        kt = -np.einsum("bij,bj->bi", Quu_inv[:, t], qu)                                 # sls.py:198
        v = (qx + np.einsum("bji,bj->bi", Qux[:, t], kt) + np.einsum("bji,bj->bi", K[:, t], qu)
             + np.einsum("bji,bj->bi", K[:, t], np.einsum("bij,bj->bi", Quu[:, t], kt)))   # sls.py:199
        k[:, t] = kt
    return k


def linear_rollout(A, Bm, K, k, dx0=None):
    """du_t = K_t dx_t + k_t ; dx_{t+1} = A_t dx_t + B_t du_t  (the closed form of l_side_inv @ rhs,
    isls.py:465, in Riccati form)."""
    Bsz, N, n, m = Bm.shape
    dx = np.zeros((Bsz, n)) if dx0 is None else dx0.copy()
    du = np.zeros((Bsz, N, m))
    dxs = np.zeros((Bsz, N, n))
    for t in range(N):
        dxs[:, t] = dx
        du[:, t] = np.einsum("bij,bj->bi", K[:, t], dx) + k[:, t]
        if t < N - 1:
            dx = np.einsum("bij,bj->bi", A[:, t], dx) + np.einsum("bij,bj->bi", Bm[:, t], du[:, t])
    return dxs, du


# ------------------------------------------------------------------------------------------ problem helpers
def _model_of(p):
    if p["model"] == "double_integrator":
        return _models.make_model("double_integrator", nb_dim=p["m"], dt=p["dt"])
    return _models.make_model(p["model"], dt=p["dt"])


def _diag_embed(d):
    out = np.zeros(d.shape + (d.shape[-1],))
    i = np.arange(d.shape[-1])
    out[..., i, i] = d
    return out


def _bounds(p, key, N, dim):
    lo, hi = p.get("lo_" + key), p.get("hi_" + key)
    if lo is None and hi is None:
        return None
    lo = np.full((N, dim), -np.inf) if lo is None else np.broadcast_to(np.asarray(lo, float), (N, dim))
    hi = np.full((N, dim), np.inf) if hi is None else np.broadcast_to(np.asarray(hi, float), (N, dim))
    return lo, hi


def initial_rollout(p):
    """What the reference user does before calling a solver (notebooks, e.g. Car/...control constraints cell 11):
    roll the initial control guess out from x0 and set it as the nominal trajectory."""
    model = _model_of(p)
    x0 = np.asarray(p["x0"], dtype=np.float64)
    B = x0.shape[0]
    u0 = np.asarray(p["u0"], dtype=np.float64)
    u_nom = np.broadcast_to(u0, (B,) + u0.shape[-2:]).copy()
    x_nom = rollout_open(model, x0, u_nom[:, None])[:, 0]
    return x_nom, u_nom


# -------------------------------------------------------------------------------------------- plain iLQR (DP)
def ilqr_dp(p, max_iter=100, L=25, tol_fun=1e-5, fixed_budget=False):
    """iSLS.solve(method='dp') (isls.py:54-132 + 336-374) for the quadratic via-point cost or the pseudo-Huber cost
    with analytic get_Cs (Cux = 0)."""
    model = _model_of(p)
    N, n, m = p["N"], p["n"], p["m"]
    x_nom, u_nom = initial_rollout(p)
    B = x_nom.shape[0]
    zs = _zs_b(p, B)
    seq = p["seq"]
    al = alphas(L)
    Rv = _R(p)
    cost = total_cost(p, zs, x_nom[:, None], u_nom[:, None])[:, 0]
    cost_log = np.full((B, max_iter + 1), np.nan)
    cost_log[:, 0] = cost
    n_log = np.ones(B, dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    alpha_idx = np.full((B, max_iter), -1, dtype=np.int32)
    active = np.ones(B, dtype=bool)
    Cuu = _diag_embed(np.broadcast_to(2.0 * Rv, (N, m)))[None]
    K_out = np.zeros((B, N, m, n))
    k_out = np.zeros((B, N, m))
    for it in range(max_iter):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        xn, un = x_nom[idx], u_nom[idx]
        A, Bm = model.get_AB(xn, un)
        cx, hxx = state_grad_hess(p, zs[idx][:, seq], xn)
        Cxx = _diag_embed(hxx)
        cu = 2.0 * Rv * un
        K, k, non_pd = backward_pass(A, Bm, cx, cu, Cxx, Cuu)
        K_out[idx], k_out[idx] = K, k
        k_cand = k[:, None] * al[None, :, None, None]                         # isls.py:357
        xs, us = rollout_closed(model, xn, un, K, k_cand)
        costs = total_cost(p, zs[idx], xs, us)
        nan = np.isnan(costs)
        costs = np.where(nan, 1e5, costs)                                     # isls.py:362
        ind = np.argmin(costs, axis=1)
        best = costs[np.arange(idx.size), ind]
        ok = (best - cost[idx]) < 0.0                                         # isls.py:365-367
        ok &= ~non_pd
        iters[idx] += 1
        alpha_idx[idx, it] = np.where(ok, ind, -1)
        acc = idx[ok]
        x_nom[acc] = xs[ok, ind[ok]]
        u_nom[acc] = us[ok, ind[ok]]
        # nominal_values setter re-evaluates the cost (isls_base.py:80-85)
        newc = total_cost(p, zs[acc], x_nom[acc][:, None], u_nom[acc][:, None])[:, 0]
        cost[acc] = newc
        cost_log[acc, n_log[acc]] = newc
        n_log[acc] += 1
        status[idx[non_pd]] |= ST_NON_PD
        status[idx[nan.any(axis=1)]] |= ST_NAN_COST
        if not fixed_budget:
            for j, b in enumerate(idx):
                nl = n_log[b]
                small = nl >= 2 and abs(cost_log[b, nl - 1] - cost_log[b, nl - 2]) < tol_fun   # isls.py:125
                if small:
                    status[b] |= ST_CONVERGED_COST
                    active[b] = False
                elif not ok[j]:                                               # isls.py:128
                    status[b] |= ST_LINESEARCH_FAIL
                    active[b] = False
    status[active] |= ST_MAX_ITER
    return dict(x=x_nom, u=u_nom, cost=cost, cost_log=cost_log, n_log=n_log, status=status, iters=iters,
                alpha_idx=alpha_idx, K=K_out, k=k_out)


# ------------------------------------------------------------------------- state projection onto obstacle sets
def project_square_batch(x, l, u):
    """isls/projections.py:246-255: rows of x onto l <= ||x||_inf <= u (only the largest component is pushed out)."""
    z = x.copy()
    j = np.argmax(np.abs(x), axis=-1)
    inside = np.max(np.abs(x), axis=-1) < l
    r = np.nonzero(inside)[0]
    z[r, j[r]] = l * np.sign(x[r, j[r]])
    return np.maximum(np.minimum(z, u), -u)


def obstacle_projections(ob):
    """The per-set projections of the parking notebook (Car/Iterative LQR with state constraints.ipynb cell 18):
    rows y[:, :2] -> centre + W^-1 Pi_square(W (y - centre)); other components untouched."""
    def make(i):
        W, Wi, c = ob["W"][i], ob["W_inv"][i], ob["centers"][i]

        def f(y):
            y_ = y.copy()
            z = y_[:, :2] - c[None]
            zp = project_square_batch(z @ W.T, ob["lower"][i], ob["upper"])
            y_[:, :2] = zp @ Wi.T + c[None]
            return y_
        return f
    return [make(i) for i in range(len(ob["centers"]))]


def project_quadratic_batch(x, l, u):
    """isls/projections.py:91-105: rows of x onto l <= 0.5 ||x||^2 <= u (radial scaling)."""
    z = x.copy()
    val = 0.5 * np.sum(x * x, axis=-1)
    c1 = np.nonzero(val > u)[0]
    c2 = np.nonzero(l > val)[0]
    z[c1] = x[c1] * np.sqrt(2 * u) / np.linalg.norm(x[c1], axis=-1)[:, None]
    z[c2] = x[c2] * np.sqrt(2 * l) / np.linalg.norm(x[c2], axis=-1)[:, None]
    return z


def sphere_projections(ob):
    """Per-obstacle projections of the spherical-obstacle notebook (Double integrator/LQR and SLS with spherical
    obstacle avoidance.ipynb cell 12): positions p -> c + Pi_quadratic(p - c, lower, upper)."""
    return [(lambda x, i=i: project_quadratic_batch(x - ob["centers"][i], ob["lower"][i], ob["upper"]) + ob["centers"][i])
            for i in range(len(ob["centers"]))]


def project_set_convex_dykstra(x0, projections, max_iter, tol):
    """isls/projections.py:465-505: Dykstra's alternating projections on all rows together; stops when every row's
    summed squared increment change is below tol, or after max_iter + 1 sweeps.  Returns (u, sweeps)."""
    d = len(projections)
    u = x0.copy()
    z = np.zeros((d,) + x0.shape)
    k = 0
    cI = np.full(x0.shape[0], 10.0)
    while k <= max_iter and np.any(cI >= tol):
        cI = cI * 0
        for i in range(d):
            prev_u = u.copy()
            u = projections[i](prev_u - z[i])
            prev_z = z[i].copy()
            z[i] = u - (prev_u - prev_z)
            cI = cI + np.linalg.norm(prev_z - z[i], axis=-1) ** 2
        k += 1
    return u, k


def project_positions_spheres(pre, ob):
    """project_state of the spherical-obstacle notebook on one problem's rows pre [N, n]: positions through
    project_set_convex (rho, max_iter, threshold) then Dykstra; the other components pass through."""
    proj = sphere_projections(ob)
    z = pre.copy()
    p1, its = project_set_convex_rows(pre[:, :2], proj, ob["rho"], ob["max_iter"], ob["threshold"])
    p2, sweeps = project_set_convex_dykstra(p1, proj, ob["dykstra_max_iter"], ob["dykstra_tol"])
    z[:, :2] = p2
    return z, its, sweeps


def project_set_convex_rows(x0, projections, rho, max_iter, threshold):
    """isls/projections.py:289-374 for As = I, bs = 0 (what both obstacle notebooks pass): consensus ADMM over the
    sets, rows of x0 [rows, dim] projected together, stop on the MAX over sets and rows of the residual norms (< threshold),
    or when both maxima change by < 1e-5 relative, or after max_iter.  Returns (x [rows, dim], iterations)."""
    K = len(projections)
    dim = x0.shape[-1]
    x = x0.T.copy()
    z = [x.copy() for _ in range(K)]
    lmb = [np.zeros_like(x) for _ in range(K)]
    l_side_inv = np.linalg.inv(np.eye(dim) + rho * K * np.eye(dim))
    prim_ = dual_ = 1e5
    it = 0
    for j in range(max_iter):
        it = j + 1
        r_side = 0.0
        for i in range(K):
            r_side = r_side + (z[i] - lmb[i])
        x = l_side_inv @ (x0.T + rho * r_side)
        z_prev = [zi for zi in z]
        pprim, pdual = prim_, dual_
        pn = np.zeros((K, x.shape[1]))
        dn = np.zeros((K, x.shape[1]))
        for i in range(K):
            z[i] = projections[i]((x + lmb[i]).T).T
            prim = x - z[i]
            dual = rho * (z[i] - z_prev[i])
            lmb[i] = lmb[i] + prim
            pn[i] = np.linalg.norm(prim, axis=0)
            dn[i] = np.linalg.norm(dual, axis=0)
        prim_, dual_ = np.max(pn), np.max(dn)
        if prim_ < threshold and dual_ < threshold:
            break
        if j < max_iter - 1:
            pc = np.abs(pprim - prim_) / (pprim + 1e-30)
            dc = np.abs(pdual - dual_) / (pdual + 1e-30)
            if pc < 1e-5 and dc < 1e-5:
                break
    return x.T, it


# ----------------------------------------------------------------------------------------------- iLQR-ADMM
def ilqr_admm(p, fixed_budget=False, outer_tol=1e-3, keep_trace=False):
    """Riccati-form restatement of iSLS.ilqr_admm (isls.py:379-501) + ADMM (admm.py:6-106) with box projections.

    p : problem dict (oracle/problems.py).  Returns final nominal (x,u), cost_log, per-problem iteration
    counts, ADMM residual logs, final z / lambda and the clip masks of the last projection.
    """
    model = _model_of(p)
    N, n, m = p["N"], p["n"], p["m"]
    I_o, I_a, L, tol, relax = p["I_o"], p["I_a"], p["L"], p["tol"], p.get("alpha", 1.0)
    x_nom, u_nom = initial_rollout(p)
    B = x_nom.shape[0]
    zs = _zs_b(p, B)
    seq = p["seq"]
    R = _R(p)                                                                 # [m] diagonal of R
    quadratic = p.get("cost", "quadratic") == "quadratic"
    bx, bu = _bounds(p, "x", N, n), _bounds(p, "u", N, m)
    obst = p.get("obstacles")
    obst_proj = obstacle_projections(obst) if obst is not None else None
    inner_log = np.zeros((B, I_o, I_a), dtype=np.int32) if obst is not None else None
    proj_x, proj_u = bx is not None or obst is not None, bu is not None
    rho_x = np.broadcast_to(np.asarray(p["rho_x"], float), (N, n)) if proj_x else np.zeros((N, n))
    rho_u = np.broadcast_to(np.asarray(p["rho_u"], float), (N, m)) if proj_u else np.zeros((N, m))
    al = alphas(L)

    cost = total_cost(p, zs, x_nom[:, None], u_nom[:, None])[:, 0]
    cost_log = np.full((B, I_o + 1), np.nan)
    cost_log[:, 0] = cost
    n_log = np.ones(B, dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    outer_iters = np.zeros(B, dtype=np.int32)
    admm_iters = np.zeros((B, I_o), dtype=np.int32)
    admm_exit = np.zeros((B, I_o), dtype=np.int32)
    res_log = np.full((B, I_o, I_a, 2), np.nan)
    alpha_idx = np.full((B, I_o, I_a), -1, dtype=np.int32)
    z_x, z_u = np.zeros((B, N, n)), np.zeros((B, N, m))
    lam_x, lam_u = np.zeros((B, N, n)), np.zeros((B, N, m))
    mask_x = np.zeros((B, N, n), dtype=np.int8)      # -1 clipped at lo, +1 clipped at hi (last projection)
    mask_u = np.zeros((B, N, m), dtype=np.int8)
    trace = [] if keep_trace else None
    active = np.ones(B, dtype=bool)

    if quadratic:
        Cxx = _diag_embed(2.0 * (np.asarray(p["Qdiag"], float)[seq] + rho_x))[None]   # recipe step 1
    Cuu = _diag_embed(2.0 * (R + rho_u))[None]
    Cuu_last = 2.0 * (R + rho_u[-1])

    for j in range(I_o):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        xn, un = x_nom[idx].copy(), u_nom[idx].copy()
        prev_cost = cost[idx].copy()
        A, Bm = model.get_AB(xn, un)                                          # isls.py:424
        zero_cx = np.zeros((1, N, n))
        zero_cu = np.zeros((1, N, m))
        zs_i = zs[idx][:, seq]                                                # [b,N,n]
        gx, hxx = state_grad_hess(p, zs_i, xn)                                # get_Cs at the nominal (isls.py:427-432)
        if not quadratic:
            Cxx = _diag_embed(hxx + 2.0 * rho_x)
        K, _, non_pd, Quu, Quu_inv, Qux = backward_pass(A, Bm, zero_cx, zero_cu, Cxx, Cuu, logs=True)
        status[idx[non_pd]] |= ST_NON_PD
        zx, zu = z_x[idx].copy(), z_u[idx].copy()                             # warm start (isls.py:489-490)
        lx, lu = np.zeros_like(zx), np.zeros_like(zu)                         # lambda reset (isls.py:414-415)
        prim = np.full(idx.size, 1e6)
        dual = np.full(idx.size, 1e6)
        x_last, u_last = xn.copy(), un.copy()
        a_active = np.ones(idx.size, dtype=bool)
        for a in range(I_a):
            ia = np.nonzero(a_active)[0]
            if ia.size == 0:
                break
            g = idx[ia]
            reg_x, reg_u = zx[ia] - lx[ia], zu[ia] - lu[ia]                   # admm.py:32-33
            # ---- f_argmin (isls.py:456-478) in Riccati form
            cx = gx[ia] + 2.0 * rho_x * (xn[ia] - reg_x)
            cu = 2.0 * R * un[ia] + 2.0 * rho_u * (un[ia] - reg_u)
            k = ff_pass(A[ia], Bm[ia], cx, cu, K[ia], Quu[ia], Quu_inv[ia], Qux[ia])
            k[:, -1] = -cu[:, -1] / Cuu_last                                  # batch-form last control
            _, du = linear_rollout(A[ia], Bm[ia], K[ia], k)
            u_cand = un[ia][:, None] + al[None, :, None, None] * du[:, None]  # isls.py:468
            x_cand = rollout_open(model, xn[ia][:, 0], u_cand)                # isls.py:469
            costs = total_cost(p, zs[g], x_cand, u_cand)                      # isls.py:470
            if proj_x:
                dxr = x_cand - reg_x[:, None]
                costs = costs + np.sum(dxr * dxr * rho_x, axis=(-1, -2))      # isls.py:471-473 (diagonal Qr)
            if proj_u:
                dur = u_cand - reg_u[:, None]
                costs = costs + np.sum(dur * dur * rho_u, axis=(-1, -2))      # isls.py:474-476
            # np.argmin returns the first NaN if any (isls.py:477) - keep numpy semantics
            ind = np.argmin(costs, axis=1)
            ar = np.arange(ia.size)
            xx, xu = x_cand[ar, ind], u_cand[ar, ind]
            alpha_idx[g, j, a] = ind
            status[g[np.isnan(costs).any(axis=1)]] |= ST_NAN_COST
            x_last[ia], u_last[ia] = xx, xu
            # ---- ADMM update (admm.py:43-69)
            pprim, pdual = prim[ia].copy(), dual[ia].copy()
            pr = np.zeros(ia.size)
            dr = np.zeros(ia.size)
            if proj_x:
                zprev = zx[ia]
                pre = relax * xx + (1.0 - relax) * zprev + lx[ia]
                if obst is not None:                                          # project_set_convex over the obstacle sets
                    znew = np.empty_like(pre)
                    for q in range(ia.size):
                        znew[q], inner_log[g[q], j, a] = project_set_convex_rows(
                            pre[q], obst_proj, obst["rho"], obst["max_iter"], obst["threshold"])
                else:
                    znew = np.clip(pre, bx[0], bx[1])                         # projections.py:7-11
                    mask_x[g] = (pre > bx[1]).astype(np.int8) - (pre < bx[0]).astype(np.int8)
                r = xx - znew
                lx[ia] = lx[ia] + r
                zx[ia] = znew
                pr = pr + np.sqrt(np.sum(r * r, axis=(-1, -2)))
                dz = znew - zprev
                dr = dr + np.sqrt(np.sum(dz * dz, axis=(-1, -2)))
            if proj_u:
                zprev = zu[ia]
                pre = relax * xu + (1.0 - relax) * zprev + lu[ia]
                znew = np.clip(pre, bu[0], bu[1])
                mask_u[g] = (pre > bu[1]).astype(np.int8) - (pre < bu[0]).astype(np.int8)
                r = xu - znew
                lu[ia] = lu[ia] + r
                zu[ia] = znew
                pr = pr + np.sqrt(np.sum(r * r, axis=(-1, -2)))
                dz = znew - zprev
                dr = dr + np.sqrt(np.sum(dz * dz, axis=(-1, -2)))
            prim[ia], dual[ia] = pr, dr
            res_log[g, j, a, 0], res_log[g, j, a, 1] = pr, dr
            admm_iters[g, j] = a + 1
            if keep_trace:
                trace.append(dict(j=j, a=a, g=g.copy(), du=du.copy(), costs=costs.copy(), ind=ind.copy(),
                                  reg_x=reg_x.copy(), reg_u=reg_u.copy(), x=xx.copy(), u=xu.copy(),
                                  z_x=zx[ia].copy(), z_u=zu[ia].copy(), lam_x=lx[ia].copy(), lam_u=lu[ia].copy()))
            if not fixed_budget:
                conv = (pr < tol) & (dr < tol)                                # admm.py:72
                pch = np.abs(pprim - pr) / (pprim + 1e-30)                    # admm.py:78-79
                dch = np.abs(pdual - dr) / (pdual + 1e-30)
                stall = (~conv) & (pch < tol) & (dch < tol)                   # admm.py:80
                admm_exit[g[conv], j] = ADMM_CONVERGED
                admm_exit[g[stall], j] = ADMM_STALLED
                a_active[ia[conv | stall]] = False
        admm_exit[idx[a_active], j] = ADMM_MAXIT
        # ---- after ADMM (isls.py:488-499)
        x_nom[idx], u_nom[idx] = x_last, u_last                              # nominal <- last primal iterate
        newc = total_cost(p, zs[idx], x_last[:, None], u_last[:, None])[:, 0]
        cost[idx] = newc
        cost_log[idx, n_log[idx]] = newc
        n_log[idx] += 1
        z_x[idx], z_u[idx] = zx, zu
        lam_x[idx], lam_u[idx] = lx, lu
        outer_iters[idx] = j + 1
        if not fixed_budget:
            for q, b in enumerate(idx):
                if abs(newc[q] - prev_cost[q]) < outer_tol:                   # isls.py:493
                    status[b] |= ST_CONVERGED_COST
                    active[b] = False
                    continue
                nl = n_log[b]
                last4 = cost_log[b, max(0, nl - 4):nl]
                prev4 = cost_log[b, max(0, nl - 8):max(0, nl - 4)]
                if prev4.size and abs(np.mean(last4) - np.mean(prev4)) < outer_tol:   # isls.py:497
                    status[b] |= ST_OSCILLATING
                    active[b] = False
    status[active] |= ST_MAX_ITER
    out = dict(x=x_nom, u=u_nom, cost=cost, cost_log=cost_log, n_log=n_log, status=status,
               outer_iters=outer_iters, admm_iters=admm_iters, admm_exit=admm_exit, res_log=res_log,
               alpha_idx=alpha_idx, z_x=z_x, z_u=z_u, lam_x=lam_x, lam_u=lam_u, mask_x=mask_x, mask_u=mask_u)
    if inner_log is not None:
        out["inner_iters"] = inner_log
    if keep_trace:
        out["trace"] = trace
    return out


# ------------------------------------------------------------------------------------- LQT-ADMM with Riccati
def lqt_admm_dp(p, fixed_budget=False, batch_form=False):
    """SLS.ADMM_LQT_DP (sls.py:298-317): linear dynamics, one Riccati pass (sls.py:85-166), then per ADMM
    iteration the feed-forward recursion (sls.py:168-202) + closed-loop linear rollout
    (sls_base.py:76-89) + projection / dual update (admm.py).  Budget p['I_a'] iterations, tolerance p['tol'].

    batch_form=True restates SLS.ADMM_LQT_Batch (sls.py:250-294) in the same Riccati form: the dense least-squares
    argmin is the same LQ minimiser except that its last control is solved for, u_{N-1} = (R + Rr)^-1 Rr reg_u, and
    ADMM is warm-started at the unconstrained solution (z_x_init, z_u_init, sls.py:266-268)."""
    model = _model_of(p)
    N, n, m = p["N"], p["n"], p["m"]
    I_a, tol, relax = p["I_a"], p["tol"], p.get("alpha", 1.0)
    x0 = np.asarray(p["x0"], dtype=np.float64)
    B = x0.shape[0]
    zs = _zs_b(p, B)
    seq = p["seq"]
    Qd = p["Qdiag"][seq]
    R = p["u_std"]
    bx, bu = _bounds(p, "x", N, n), _bounds(p, "u", N, m)
    obst = p.get("obstacles")
    proj_x, proj_u = bx is not None or obst is not None, bu is not None
    inner_log = np.zeros((B, I_a, 2), dtype=np.int32) if obst is not None else None
    rho_x = np.broadcast_to(np.asarray(p["rho_x"], float), (N, n)) if proj_x else np.zeros((N, n))
    rho_u = np.broadcast_to(np.asarray(p["rho_u"], float), (N, m)) if proj_u else np.zeros((N, m))
    A1, B1 = model.A, model.B
    A = np.broadcast_to(A1, (1, N, n, n))
    Bm = np.broadcast_to(B1, (1, N, n, m))
    Cxx = _diag_embed(2.0 * (Qd + rho_x))[None]
    Cuu = _diag_embed(2.0 * (R + rho_u))[None]
    K, _, non_pd, Quu, Quu_inv, Qux = backward_pass(A, Bm, np.zeros((1, N, n)), np.zeros((1, N, m)), Cxx, Cuu,
                                                    logs=True)
    Ab = np.broadcast_to(A, (B, N, n, n))
    Bb = np.broadcast_to(Bm, (B, N, n, m))
    Kb = np.broadcast_to(K, (B, N, m, n))
    Quub, Quuib, Quxb = (np.broadcast_to(a, (B,) + a.shape[1:]) for a in (Quu, Quu_inv, Qux))
    zs_i = zs[:, seq]
    zx, zu = np.zeros((B, N, n)), np.zeros((B, N, m))
    if batch_form:                                                             # sls.py:266-268
        K0, _, _, Quu0, Quui0, Qux0 = backward_pass(A, Bm, np.zeros((1, N, n)), np.zeros((1, N, m)),
                                                    _diag_embed(2.0 * Qd)[None],
                                                    _diag_embed(np.full((N, m), 2.0 * R))[None], logs=True)
        bc = lambda a: np.broadcast_to(a, (B,) + a.shape[1:])
        k0 = ff_pass(Ab, Bb, -2.0 * Qd * zs_i, np.zeros((B, N, m)), bc(K0), bc(Quu0), bc(Quui0), bc(Qux0))
        zx, zu = linear_rollout(Ab, Bb, bc(K0), k0, dx0=x0)
    lx, lu = np.zeros((B, N, n)), np.zeros((B, N, m))
    x_last, u_last = np.zeros((B, N, n)), np.zeros((B, N, m))
    k_last = np.zeros((B, N, m))
    prim, dual = np.full(B, 1e6), np.full(B, 1e6)
    res_log = np.full((B, I_a, 2), np.nan)
    iters = np.zeros(B, dtype=np.int32)
    exit_code = np.zeros(B, dtype=np.int32)
    mask_x = np.zeros((B, N, n), dtype=np.int8)
    mask_u = np.zeros((B, N, m), dtype=np.int8)
    a_active = np.ones(B, dtype=bool)
    for a in range(I_a):
        ia = np.nonzero(a_active)[0]
        if ia.size == 0:
            break
        reg_x, reg_u = zx[ia] - lx[ia], zu[ia] - lu[ia]
        cx = -2.0 * Qd * zs_i[ia] - 2.0 * rho_x * reg_x                        # sls.py:187-193
        cu = -2.0 * rho_u * reg_u
        k = ff_pass(Ab[ia], Bb[ia], cx, cu, Kb[ia], Quub[ia], Quuib[ia], Quxb[ia])
        if batch_form:
            k[:, -1] = -cu[:, -1] / (2.0 * (R + rho_u[-1]))
        xs, us = linear_rollout(Ab[ia], Bb[ia], Kb[ia], k, dx0=x0[ia])         # sls_base.py:76-89
        x_last[ia], u_last[ia], k_last[ia] = xs, us, k
        pprim, pdual = prim[ia].copy(), dual[ia].copy()
        pr, dr = np.zeros(ia.size), np.zeros(ia.size)
        if proj_x:
            zprev = zx[ia]
            pre = relax * xs + (1.0 - relax) * zprev + lx[ia]
            if obst is not None:
                znew = np.empty_like(pre)
                for q in range(ia.size):
                    znew[q], i1, i2 = project_positions_spheres(pre[q], obst)
                    inner_log[ia[q], a] = i1, i2
            else:
                znew = np.clip(pre, bx[0], bx[1])
                mask_x[ia] = (pre > bx[1]).astype(np.int8) - (pre < bx[0]).astype(np.int8)
            r = xs - znew
            lx[ia] = lx[ia] + r
            zx[ia] = znew
            pr = pr + np.sqrt(np.sum(r * r, axis=(-1, -2)))
            dz = znew - zprev
            dr = dr + np.sqrt(np.sum(dz * dz, axis=(-1, -2)))
        if proj_u:
            zprev = zu[ia]
            pre = relax * us + (1.0 - relax) * zprev + lu[ia]
            znew = np.clip(pre, bu[0], bu[1])
            mask_u[ia] = (pre > bu[1]).astype(np.int8) - (pre < bu[0]).astype(np.int8)
            r = us - znew
            lu[ia] = lu[ia] + r
            zu[ia] = znew
            pr = pr + np.sqrt(np.sum(r * r, axis=(-1, -2)))
            dz = znew - zprev
            dr = dr + np.sqrt(np.sum(dz * dz, axis=(-1, -2)))
        prim[ia], dual[ia] = pr, dr
        res_log[ia, a, 0], res_log[ia, a, 1] = pr, dr
        iters[ia] = a + 1
        if not fixed_budget:
            conv = (pr < tol) & (dr < tol)
            pch = np.abs(pprim - pr) / (pprim + 1e-30)
            dch = np.abs(pdual - dr) / (pdual + 1e-30)
            stall = (~conv) & (pch < tol) & (dch < tol)
            exit_code[ia[conv]] = ADMM_CONVERGED
            exit_code[ia[stall]] = ADMM_STALLED
            a_active[ia[conv | stall]] = False
    exit_code[a_active] = ADMM_MAXIT
    cost = quad_cost(p, zs, x_last[:, None], u_last[:, None])[:, 0]
    return dict(x=x_last, u=u_last, K=np.broadcast_to(K, (B, N, m, n)).copy(), k=k_last, cost=cost, iters=iters,
                exit_code=exit_code, res_log=res_log, z_x=zx, z_u=zu, lam_x=lx, lam_u=lu, mask_x=mask_x,
                mask_u=mask_u, non_pd=np.broadcast_to(non_pd, (B,)).copy(), inner_iters=inner_log)


# ================================================================================ SLS path (item 4, config 4)
def build_Sw_Su(A, Bm, N):
    """Base.AB setter (isls/base.py:98-119): Sw = (I - Z A)^-1 (lower block triangular, blocks A^(i-j)),
    Su blocks A^(i-j-1) B for i > j.  Built with the reference's backward block recursion."""
    n, m = Bm.shape
    Sw = np.kron(np.triu(np.ones((N, N))).T, np.eye(n))           # base.py:18
    Su = np.zeros((n * N, N * m))                                 # base.py:19
    for i in range(N - 1, 0, -1):
        Su[i * n:, (i - 1) * m:i * m] = Sw[i * n:, i * n:(i + 1) * n] @ Bm       # base.py:115-116
        Sw[i * n:, (i - 1) * n:i * n] = Sw[i * n:, i * n:(i + 1) * n] @ A        # base.py:118-119
    return Sw, Su


def trailing_inverses(L, m, N):
    """Base.compute_inverses (isls/base.py:44-50): inverses of all trailing principal sub-matrices L[i m:, i m:],
    i = 0..N, each obtained from the previous by the rank-2m Woodbury down-date of isls/base.py:32-42, which is the
    Schur-complement identity  D^-1 = S - r p^-1 q  for  inv([[a,b],[c,D]]) = [[p,q],[r,S]]."""
    invs = [np.linalg.inv(L)]
    for i in range(N):
        Ai = invs[i]
        p, q, r, S = Ai[:m, :m], Ai[:m, m:], Ai[m:, :m], Ai[m:, m:]
        invs.append(S - r @ np.linalg.solve(p, q))
    return invs


def project_soc_unit_batch(z, t):
    """isls/projections.py:140-162 verbatim semantics (incl. SURVEY D9: every row with t < 0 is zeroed)."""
    z_norm = np.linalg.norm(z, axis=-1)
    z_ = z.copy()
    t_ = t.copy()
    cond1 = np.logical_or(z_norm <= -t, t < 0)
    cond2 = np.logical_or(z_norm > t, z_norm > -t)
    cond3 = z_norm <= t
    tmp = (z_norm + t_) / 2
    z_[cond2] = tmp[cond2, None] * z[cond2] / (z_norm[cond2, None] + 1e-30)
    t_[cond2] = tmp[cond2].copy()
    z_[cond1] = 0.
    t_[cond1] = 0.
    z_[cond3] = z[cond3]
    t_[cond3] = t[cond3]
    return z_, t_


def project_set_convex_soc(x0, As, bs, rho=1.0, max_iter=200, threshold=1e-4):
    """isls/projections.py:289-374 with projections = [project_soc_unit] * len(As): inner ADMM onto the
    intersection of {A_i x + b_i in SOC}; x0 [rows, c].  Returns (x [rows, c], inner iterations)."""
    P = len(As)
    x0 = np.asarray(x0)
    rows = x0.shape[0]
    x = x0.T.copy()
    z = [As[i] @ x + bs[i][:, None] for i in range(P)]
    lmb = [zz * 0 for zz in z]
    l_side_add = sum(As[i].T @ As[i] for i in range(P))
    l_side_inv = np.linalg.inv(np.eye(x0.shape[-1]) + rho * l_side_add)
    prim_n = np.ones((P, rows)) * 1e5
    dual_n = np.ones((P, rows)) * 1e5
    prim_, dual_ = 1e5, 1e5
    its = 0
    for j in range(max_iter):
        its = j + 1
        r_side = sum(As[i].T @ (-bs[i][:, None] + z[i] - lmb[i]) for i in range(P))
        x = l_side_inv @ (x0.T + rho * r_side)
        z_prev = list(z)
        pprim, pdual = prim_, dual_
        for i in range(P):
            Ax_b = As[i] @ x + bs[i][:, None]
            y = (Ax_b + lmb[i]).T
            zz, tt = project_soc_unit_batch(y[:, :-1], y[:, -1])
            z[i] = np.concatenate([zz, tt[:, None]], axis=1).T
            prim_res = Ax_b - z[i]
            dual_res = rho * As[i].T @ (z[i] - z_prev[i])
            lmb[i] = lmb[i] + prim_res
            prim_n[i] = np.linalg.norm(prim_res, axis=0)
            dual_n[i] = np.linalg.norm(dual_res, axis=0)
        prim_, dual_ = np.max(prim_n), np.max(dual_n)
        if prim_ < threshold and dual_ < threshold:
            break
        if j < max_iter - 1:
            pch = np.abs(pprim - prim_) / (pprim + 1e-30)
            dch = np.abs(pdual - dual_) / (pdual + 1e-30)
            if pch < 1e-5 and dch < 1e-5:
                break
    return x.T, its


def sls_solve(A, Bm, N, Qdiag_t, xd, u_std):
    """SLS.solve_sls (isls/sls.py:205-233).  Qdiag_t [N, n] per-step diagonal of Q, xd [N n] (one problem) or
    [B, N n].  Returns PHI_U [N m, N n] (shared), du [B, N m], plus the operators reused by admm_sls."""
    n, m = Bm.shape
    Sw, Su = build_Sw_Su(A, Bm, N)
    q = Qdiag_t.reshape(-1)
    DTQ = Su.T * q[None, :]                                        # Su' Q with block-diagonal (here diagonal) Q
    L = DTQ @ Su + u_std * np.eye(N * m)
    invs = trailing_inverses(L, m, N)
    xd2 = np.atleast_2d(xd)
    du = (invs[0] @ DTQ @ xd2.T).T                                 # sls.py:221
    r_side = -DTQ @ Sw                                             # sls.py:225
    PHI_U = np.zeros((N * m, N * n))
    for i in range(N):
        PHI_U[i * m:, i * n:(i + 1) * n] = invs[i] @ r_side[i * m:, i * n:(i + 1) * n]     # sls.py:228-229
    return dict(PHI_U=PHI_U, du=du, Sw=Sw, Su=Su, DTQ=DTQ, L=L)


def admm_sls(A, Bm, N, Qdiag_t, xd, u_std, As, bs, rho_u, max_iter=5000, alpha=1.0, tol=1e-3, inner_rho=1.0,
             inner_max_iter=200, inner_threshold=1e-4, fixed_budget=False, x_rows=None, rho_x=None):
    """SLS.ADMM_SLS (isls/sls.py:319-454) with project_u = project_set_convex(.., [project_soc_unit]*P) row-wise.
    One problem per row of xd [B, N n].  Returns du [B, N m], phi_u [B, N m, N n], logs.

    State side (`LQR and SLS with state bounds.ipynb` cells 16-17): rho_x [N n] = diagonal of Qr, x_rows = list of
    (row index into N n, As_i, bs_i): project_x projects each listed row of [d_x | Phi_x(:, :n/2)] by its own
    project_set_convex call (same inner parameters) and leaves the other rows alone."""
    n, m = Bm.shape
    c = n // 2 + 1
    base = sls_solve(A, Bm, N, Qdiag_t, xd, u_std)
    Sw, Su, DTQ = base["Sw"], base["Su"], base["DTQ"]
    Sx = Sw[:, :n // 2]
    l_side = base["L"] + rho_u * np.eye(N * m)                     # sls.py:339-349 (Rr = rho_u I)
    r_fb = -DTQ @ Sx
    proj_x = x_rows is not None
    if proj_x:
        qr = np.asarray(rho_x, float).reshape(-1)
        SuTQr = Su.T * qr[None, :]                                 # sls.py:342-347
        l_side = l_side + SuTQr @ Su
        r_fb = r_fb - SuTQr @ Sx
    l_inv = trailing_inverses(l_side, m, N)[0]                     # sls.py:352, 367
    xd2 = np.atleast_2d(xd)
    B = xd2.shape[0]
    du_out = np.zeros((B, N * m))
    phi_out = np.zeros((B, N * m, N * n))
    logs, iters, exits, inner_total = [], np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int64)
    for b in range(B):
        r_side = np.concatenate([(DTQ @ xd2[b])[:, None], r_fb], axis=-1)
        z_u = np.zeros((N * m, c))
        lmb = np.zeros((N * m, c))
        z_x = np.zeros((N * n, c))
        lmb_x = np.zeros((N * n, c))
        prim = dual = 1e6
        lg = []
        ex = ADMM_MAXIT
        for j in range(max_iter):
            reg_u = z_u - lmb
            rs = r_side + rho_u * reg_u
            if proj_x:
                rs = rs + SuTQr @ (z_x - lmb_x)
            x_u = l_inv @ rs                                       # sls.py:372-380
            pprim, pdual = prim, dual
            if proj_x:
                x_x = Su @ x_u
                x_x[:, 1:] += Sx                                   # sls.py:381-382
                zx_prev = z_x.copy()
                y = alpha * x_x + (1 - alpha) * z_x + lmb_x
                z_x = y.copy()
                for (row, As_x, bs_x) in x_rows:
                    z_x[row:row + 1], it_in = project_set_convex_soc(y[row:row + 1], As_x, bs_x, rho=inner_rho,
                                                                     max_iter=inner_max_iter, threshold=inner_threshold)
                    inner_total[b] += it_in
                pr_x = x_x - z_x
                lmb_x = lmb_x + pr_x
            z_prev = z_u.copy()
            z_u, it_in = project_set_convex_soc(alpha * x_u + (1 - alpha) * z_u + lmb, As, bs, rho=inner_rho,
                                                max_iter=inner_max_iter, threshold=inner_threshold)
            inner_total[b] += it_in
            pr = x_u - z_u
            lmb = lmb + pr
            dual = np.linalg.norm(rho_u * (z_u - z_prev))          # sls.py:417 (Frobenius, Rr = rho_u I)
            prim = np.linalg.norm(rho_u * pr)
            if proj_x:                                             # sls.py:414-415
                dual = np.linalg.norm(qr[:, None] * (z_x - zx_prev)) + dual
                prim = np.linalg.norm(qr[:, None] * pr_x) + prim
            lg.append((prim, dual))
            iters[b] = j + 1
            if fixed_budget:
                continue
            if prim < tol and dual < tol:
                ex = ADMM_CONVERGED
                break
            pch = abs(pprim - prim) / (pprim + 1e-30)
            dch = abs(pdual - dual) / (pdual + 1e-30)
            if pch < 1e-2 and dch < 1e-2:                          # sls.py:429
                ex = ADMM_STALLED
                break
        exits[b] = ex
        du_out[b] = x_u[:, 0]
        phi_out[b] = np.concatenate([x_u[:, 1:c], base["PHI_U"][:, c - 1:]], axis=-1)   # sls.py:449-450
        logs.append(np.array(lg))
    return dict(du=du_out, phi_u=phi_out, logs=logs, iters=iters, exit_code=exits, inner_total=inner_total,
                PHI_U=base["PHI_U"], du0=base["du"], Sw=Sw, Su=Su)


def sls_controller(Sw, Su, PHI_U, du):
    """SLS.controller (isls/sls.py:235-242): PHI_X = Sw + Su PHI_U; K = PHI_U PHI_X^-1; k = (I - K Su) du."""
    PHI_X = Sw + Su @ PHI_U
    K = PHI_U @ np.linalg.inv(PHI_X)
    k = (np.eye(Su.shape[-1]) - K @ Su) @ du
    return K, k


# ============================================================================ Monte-Carlo closed-loop evaluation
def mc_rollout(model, mode, x0, K, k, N, x_nom=None, u_nom=None):
    """Noise-free restatement of get_trajectory_batch / dp / sls (isls/sls_base.py:62-105, isls/isls_base.py:28-71).
    x0 [B, n]; returns x [B, N, n], u [B, N, m]."""
    B, n = x0.shape
    m = model.m
    xs, us = np.zeros((B, N, n)), np.zeros((B, N, m))
    x = x0.copy()
    xv = np.zeros((B, N * n))
    for t in range(N):
        if mode == "batch":
            u = np.broadcast_to(k[t], (B, m))
        elif mode == "dp":
            u = x @ K[t].T + k[t]
        else:
            xv[:, t * n:(t + 1) * n] = x - (0.0 if x_nom is None else x_nom[t])
            u = (xv @ K.T + k)[:, t * m:(t + 1) * m] + (0.0 if u_nom is None else u_nom[t])
        xs[:, t], us[:, t] = x, u
        x = model.f(x, u)
    return xs, us


# ------------------------------------------------------------------------ robust nonlinear iSLS-ADMM
def isls_admm(p, fixed_budget=False):
    """iSLS.isls_admm (isls/isls.py:503-712) with project_u = project_set_convex over SOC chance constraints
    (3DoF robot/State bounds and robust control bounds.ipynb cells 24-26), in Riccati form.

    The reference solves  [d_u | Phi_u] = l_side^-1 (r_side + Rr reg_u)  with the dense (N m)^2 inverse
    (isls.py:562-579).  Column by column that is an LQ problem around the nominal trajectory with the gains K_t of one
    Riccati pass (Cxx = 2Q, Cuu = 2(R + Rr)):
      column 0 (d_u):        dx_0 = 0,    cx = 2Q(x^ - z_via), cu = 2R u^ - 2Rr reg_0
      column c >= 1 (Phi_u): dx_0 = e_c   (Sx = C[:, :dim], isls.py:546), cx = 0, cu = -2Rr reg_c
    each a feed-forward pass (sls.py:168-202) + linear rollout du_t = K_t dx_t + k_t, with the batch-form last control
    du_{N-1} = -Cuu^-1 cu.  Line search on column 0 only, cost_function without penalty terms (isls.py:586-599);
    ADMM on the matrix variable with the row-wise projection, residuals weighted by Rr (isls.py:641-654), stall
    threshold 1e-3 (isls.py:664), outer stop |dcost| < 1e-4 or oscillation (isls.py:700-706).
    State side (project_x, isls.py:556-559, 571-572, 631-638, 648-650), when p["robust"]["x"] is given: Qr = diag(rho_x)
    joins Cxx, the columns get the linear state terms cx -= 2 Qr reg_x (column 0) / cx = -2 Qr reg_x (columns >= 1),
    x_x = [x_win - x^ | dx of the linear rollouts], z_x = project_x(alpha x_x + (1 - alpha) z_x + lambda_x, x_nom) with
    the closure of ref_shim.run_isls_admm (column 0 shifted by x_nom; one project_set_convex call per listed state
    component over its N rows; all other rows pass through), residuals ||Qr .||_F added to the control side's.
    Returns per problem: x, u, cost_log, d_u [N,m], phi_u [N,m,dim], iteration counts."""
    model = _model_of(p)
    N, n, m = p["N"], p["n"], p["m"]
    rb = p["robust"]
    dim, rho = rb["dim"], float(rb["rho_u"])
    projected = rb.get("As") is not None and not rb.get("u_unprojected", False)   # no project_u: Rr = 0, z = x (cell 23)
    if not projected:
        rho = 0.0
    rx = rb.get("x")                                       # state side: dict(comps, bs [G][P][ra], rho_x [N, n])
    rhx = np.asarray(rx["rho_x"], float) if rx else np.zeros((p["N"], p["n"]))
    I_o, I_a, L, tol, relax = p["I_o"], p["I_a"], p["L"], p["tol"], p.get("alpha", 1.0)
    x_nom, u_nom = initial_rollout(p)
    B = x_nom.shape[0]
    zs = _zs_b(p, B)
    seq = p["seq"]
    Qd = np.asarray(p["Qdiag"], float)[seq]
    R = _R(p)
    al = alphas(L)
    C = dim + 1
    cost = total_cost(p, zs, x_nom[:, None], u_nom[:, None])[:, 0]
    cost_log = np.full((B, I_o + 1), np.nan)
    cost_log[:, 0] = cost
    n_log = np.ones(B, dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    outer_iters = np.zeros(B, dtype=np.int32)
    admm_iters = np.zeros((B, I_o), dtype=np.int32)
    inner_iters = np.zeros((B, I_o, I_a), dtype=np.int32)
    inner_iters_x = np.zeros((B, I_o, I_a, len(rx["comps"]) if rx else 0), dtype=np.int32)
    d_x = np.zeros((B, N, n))
    phi_x = np.zeros((B, N, n, dim))
    res_log = np.full((B, I_o, I_a, 2), np.nan)
    alpha_idx = np.full((B, I_o, I_a), -1, dtype=np.int32)
    d_u = np.zeros((B, N, m))
    phi_u = np.zeros((B, N, m, dim))
    Cxx = _diag_embed(2.0 * (Qd + rhx))[None]
    Cuu = _diag_embed(np.broadcast_to(2.0 * (R + rho), (N, m)))[None]
    Cuu_last = 2.0 * (R + rho)
    for b in range(B):
        z_u = np.zeros((N, m, C))                                             # z_u_init (isls.py:537)
        z_x = np.zeros((N, n, C))                                             # z_x_init (isls.py:536)
        xn, un = x_nom[b:b + 1].copy(), u_nom[b:b + 1].copy()
        for k in range(I_o):
            prev_cost = cost[b]
            A, Bm = model.get_AB(xn, un)
            Kg, _, non_pd, Quu, Quu_inv, Qux = backward_pass(A, Bm, np.zeros((1, N, n)), np.zeros((1, N, m)), Cxx, Cuu,
                                                             logs=True, joseph=True)
            if non_pd[0]:
                status[b] |= ST_NON_PD
            lam = np.zeros_like(z_u)                                          # lmb_u = 0 (isls.py:615)
            lam_x = np.zeros_like(z_x)                                        # lmb_x = 0 (isls.py:614)
            x_x = np.zeros((N, n, C))
            prim = dual = 1e6
            cx0 = 2.0 * Qd * (xn - zs[b:b + 1][:, seq])
            x_u = np.zeros((N, m, C))
            x_win = xn.copy()
            for j in range(I_a):
                reg = z_u - lam                                               # isls.py:624
                reg_x = z_x - lam_x                                           # isls.py:623
                # ---- f_argmin (isls.py:568-608)
                du_ = np.zeros((N, m, C))
                dx_ = np.zeros((N, n, C))
                for c in range(C):
                    if c == 0:
                        cx, cu = cx0 - 2.0 * rhx * reg_x[None, :, :, 0], 2.0 * R * un - 2.0 * rho * reg[None, :, :, 0]
                        dx0 = None
                    else:
                        cx, cu = -2.0 * rhx * reg_x[None, :, :, c], -2.0 * rho * reg[None, :, :, c]
                        dx0 = np.zeros((1, n))
                        dx0[0, c - 1] = 1.0
                    kk = ff_pass(A, Bm, cx, cu, Kg, Quu, Quu_inv, Qux)
                    kk[:, -1] = -cu[:, -1] / Cuu_last
                    dxc, duc = linear_rollout(A, Bm, Kg, kk, dx0)
                    du_[:, :, c] = duc[0]
                    dx_[:, :, c] = dxc[0]
                u_cand = un[:, None] + al[None, :, None, None] * du_[None, None, :, :, 0]
                x_cand = rollout_open(model, xn[:, 0], u_cand)
                costs = total_cost(p, zs[b:b + 1], x_cand, u_cand)            # isls.py:586 (no penalty terms)
                ind = int(np.argmin(costs[0]))
                alpha_idx[b, k, j] = ind
                x_u = du_.copy()
                x_u[:, :, 0] = al[ind] * du_[:, :, 0]                         # isls.py:602-603
                x_win = x_cand[:, ind]
                x_x = dx_.copy()
                x_x[:, :, 0] = (x_win - xn)[0]                                # isls.py:605-606
                # ---- ADMM update (isls.py:628-654)
                z_prev = z_u
                y = relax * x_u + (1.0 - relax) * z_u + lam
                y2 = y.reshape(N * m, C).copy()
                y2[:, 0] += un[0].reshape(-1)                                 # project_u(z, u_nom): notebook cell 25
                if projected:
                    zp, its = project_set_convex_soc(y2, rb["As"], rb["bs"], rho=rb["inner_rho"],
                                                     max_iter=rb["inner_max_iter"], threshold=rb["inner_threshold"])
                else:
                    zp, its = y2, 1
                zp = zp.copy()
                zp[:, 0] -= un[0].reshape(-1)
                z_u = zp.reshape(N, m, C)
                inner_iters[b, k, j] = its
                r = x_u - z_u
                lam = lam + r
                pprim, pdual = prim, dual
                dual = np.linalg.norm(rho * (z_u - z_prev).reshape(N * m, C))
                prim = np.linalg.norm(rho * r.reshape(N * m, C))
                if rx:                                                        # isls.py:631-638, 648-650
                    zx_prev = z_x
                    yx = (relax * x_x + (1.0 - relax) * z_x + lam_x).reshape(N * n, C).copy()
                    yx[:, 0] += xn[0].reshape(-1)
                    for g, comp in enumerate(rx["comps"]):
                        rows = np.arange(N) * n + comp
                        yx[rows], itx = project_set_convex_soc(yx[rows], rb["As"], rx["bs"][g], rho=rb["inner_rho"],
                                                               max_iter=rb["inner_max_iter"],
                                                               threshold=rb["inner_threshold"])
                        inner_iters_x[b, k, j, g] = itx
                    yx[:, 0] -= xn[0].reshape(-1)
                    z_x = yx.reshape(N, n, C)
                    r_x = x_x - z_x
                    lam_x = lam_x + r_x
                    dual = np.linalg.norm((rhx[:, :, None] * (z_x - zx_prev)).reshape(N * n, C)) + dual
                    prim = np.linalg.norm((rhx[:, :, None] * r_x).reshape(N * n, C)) + prim
                res_log[b, k, j] = prim, dual
                admm_iters[b, k] = j + 1
                if not fixed_budget:
                    if prim < tol and dual < tol:
                        break
                    pc = abs(pprim - prim) / (pprim + 1e-30)
                    dc = abs(pdual - dual) / (pdual + 1e-30)
                    if pc < 1e-3 and dc < 1e-3:
                        break
            # ---- new nominal (isls.py:690-693)
            un = un + x_u[None, :, :, 0]
            xn = xn + (x_win - xn)
            newc = total_cost(p, zs[b:b + 1], xn[:, None], un[:, None])[0, 0]
            cost[b] = newc
            cost_log[b, n_log[b]] = newc
            n_log[b] += 1
            outer_iters[b] = k + 1
            d_u[b] = x_u[:, :, 0]
            phi_u[b] = x_u[:, :, 1:]
            d_x[b] = x_x[:, :, 0]
            phi_x[b] = x_x[:, :, 1:]
            if not fixed_budget:
                if abs(newc - prev_cost) < 1e-4:                              # isls.py:700
                    status[b] |= ST_CONVERGED_COST
                    break
                nl = n_log[b]
                last4 = cost_log[b, max(0, nl - 4):nl]
                prev4 = cost_log[b, max(0, nl - 8):max(0, nl - 4)]
                if prev4.size and abs(np.mean(last4) - np.mean(prev4)) < 1e-3:  # isls.py:704
                    status[b] |= ST_OSCILLATING
                    break
        else:
            status[b] |= ST_MAX_ITER
        x_nom[b], u_nom[b] = xn[0], un[0]
    return dict(x=x_nom, u=u_nom, cost=cost, cost_log=cost_log, n_log=n_log, status=status, outer_iters=outer_iters,
                admm_iters=admm_iters, inner_iters=inner_iters, res_log=res_log, alpha_idx=alpha_idx, d_u=d_u,
                phi_u=phi_u, d_x=d_x, phi_x=phi_x, inner_iters_x=inner_iters_x)
