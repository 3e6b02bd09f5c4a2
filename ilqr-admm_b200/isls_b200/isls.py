"""iSLS - batched, device-resident mirror of the reference's `isls.iSLS` hot path (isls/isls.py, isls/isls_base.py,
isls/base.py).  Same constructor, properties and method names/keywords (HEAD and the legacy spellings used by the
notebooks and README), plus an optional leading problem axis B on every array.  All arithmetic runs in
libisls_b200.so; there is no CPU fallback.
"""
import numpy as np
import torch

from . import solver as S
from ._lib import IslsError
from .projections import Bound, ObstacleSets, SetConvexSOC, SetConvexSOCComponents
from .utils import diag_of

_MODEL_NAMES = ("car", "arm3", "double_integrator", "tassa_car")


class PseudoHuberCost:
    """Device descriptor of the Tutorial's cost closure (notebooks/Tutorial.ipynb cell 14):
        sum_t [ sum_j cu_j u_tj^2 + sum_i cx_i ph(x_ti, px_i) ] + sum_i cf_i ph(x_{N-1,i}, pf_i),
        ph(x, p) = sqrt(x^2 + p^2) - p.
    cx / px may be shorter than x_dim (the notebook's running cost acts on (x, y) only); assign it to
    `iSLS.cost_function`.  Its analytic derivatives (the autograd `get_Cs` of cell 16) are evaluated by the kernels."""

    def __init__(self, cu, cx, px, cf, pf):
        self.cu, self.cx, self.px, self.cf, self.pf = (np.asarray(a, dtype=np.float64).reshape(-1)
                                                       for a in (cu, cx, px, cf, pf))

    def arrays(self, n, N):
        def pad(a, fill):
            out = np.full(n, fill)
            out[:a.size] = a
            return out
        cx, px = pad(self.cx, 0.0), pad(self.px, 1.0)
        seq = np.zeros(N, dtype=np.int32)
        seq[-1] = 1
        return dict(zs=np.zeros((2, n)), seq=seq, Ws=np.stack([cx, cx]), Ps=np.stack([px, px]),
                    Ws_b=np.stack([np.zeros(n), pad(self.cf, 0.0)]), Ps_b=np.stack([np.ones(n), pad(self.pf, 1.0)]),
                    Rdiag=self.cu)


class iSLS:
    def __init__(self, x_dim, u_dim, N, batch=None, device="cuda:0"):
        """iSLS(x_dim, u_dim, N) (isls/isls.py:9,52); `batch` adds the problem axis (None = one problem)."""
        self.x_dim, self.u_dim, self.N = int(x_dim), int(u_dim), int(N)
        self.batch = batch
        self.nb = 1 if batch is None else int(batch)
        self.device = device
        self.alphas = S.alphas(50)                       # isls_base.py:10-11
        self._model = None
        self._model_kw = {}
        self.A = self.B = None
        self.reset()
        self.zs = self.Qs = self.seq = self.Rt = None
        self._cost, self._cost_kw = "quadratic", {}
        self._plan_cache = {}

    # ------------------------------------------------------------------ plugin slots
    @property
    def forward_model(self):
        return self._model

    @forward_model.setter
    def forward_model(self, spec):
        """Name of a registered device model ("car", "arm3", "double_integrator"), optionally (name, {"dt": ..})
        (slot: isls/isls_base.py:106-111)."""
        if isinstance(spec, (tuple, list)):
            name, kw = spec[0], dict(spec[1])
        else:
            name, kw = spec, {}
        if callable(name) and hasattr(name, "isls_model"):
            name, kw = name.isls_model, dict(getattr(name, "isls_model_kw", {}))
        if not isinstance(name, str) or name not in _MODEL_NAMES:
            raise TypeError("forward_model must name a registered device model %s; Python callables cannot run "
                            "inside the kernels and there is no CPU fallback" % (_MODEL_NAMES,))
        self._model, self._model_kw = name, kw
        self._plan_cache.clear()

    @property
    def cost_function(self):
        return "pseudo_huber" if self._cost == "pseudo_huber" else "quadratic_viapoint"

    @cost_function.setter
    def cost_function(self, f):
        """Slot isls/isls_base.py:113-131.  Device costs: the quadratic via-point cost (set_quadratic_cost) and the
        Tutorial's pseudo-Huber cost (a `PseudoHuberCost` descriptor, or set_pseudo_huber_cost)."""
        if isinstance(f, PseudoHuberCost):
            self.set_pseudo_huber_cost(**f.arrays(self.x_dim, self.N))
        elif f not in (None, "quadratic_viapoint"):
            raise TypeError("cost_function must be None / 'quadratic_viapoint' (set_quadratic_cost) or a "
                            "PseudoHuberCost descriptor; Python callables cannot run inside the kernels")

    def _check_get_AB(self, get_AB):
        if get_AB is None:
            return
        name = getattr(get_AB, "isls_model", get_AB)
        if name != self._model:
            raise TypeError("get_AB must be None or the name of the device model set as forward_model (%r); "
                            "Jacobians are evaluated by the device model" % (self._model,))

    def _check_get_Cs(self, get_Cs):
        """The cost derivatives (isls.py:263-279, Tutorial cell 16) are evaluated analytically by the kernels for the
        cost set on this object; a Python get_Cs callable cannot run there."""
        if get_Cs is not None and get_Cs not in ("analytic", self.cost_function):
            raise TypeError("get_Cs must be None or 'analytic': the device evaluates the derivatives of the cost "
                            "set through set_quadratic_cost / cost_function")

    # ------------------------------------------------------------------ cost
    def set_quadratic_cost(self, zs, Qs, seq, u_std):
        """isls/base.py:81-89.  zs [k,n] or [B,k,n]; Qs [k,n,n] (diagonal) or [k,n] diagonals."""
        zs = np.asarray(zs, dtype=np.float64)
        Qs = np.asarray(Qs, dtype=np.float64)
        self.zs = zs
        self.Qs = Qs
        self.Qdiag = diag_of(Qs, "Qs") if Qs.ndim == 3 else Qs
        self.seq = np.asarray(seq, dtype=np.int32)
        self.u_std = float(u_std)
        self.Rt = np.eye(self.u_dim) * u_std
        self._cost, self._cost_kw = "quadratic", {}
        self._plan_cache.clear()

    set_cost_variables = set_quadratic_cost          # legacy name (README.md:24-39)

    def set_pseudo_huber_cost(self, zs, Ws, Ps, seq, Rdiag, Ws_b=None, Ps_b=None):
        """Pseudo-Huber state cost sum_t sum_i W[seq_t, i] (sqrt((x_ti - z[seq_t, i])^2 + P[seq_t, i]^2) - P[seq_t, i])
        (+ the same with Ws_b, Ps_b) + sum_t u' diag(Rdiag) u: notebooks/Tutorial.ipynb cell 14 in the (zs, seq)
        via-point form of set_quadratic_cost."""
        self.zs = np.asarray(zs, dtype=np.float64)
        self.Qdiag = np.asarray(Ws, dtype=np.float64)
        self.Qs = self.Qdiag
        self.seq = np.asarray(seq, dtype=np.int32)
        Rdiag = np.asarray(Rdiag, dtype=np.float64).reshape(self.u_dim)
        self.u_std = float(Rdiag[0])
        self.Rt = np.diag(Rdiag)
        self._cost = "pseudo_huber"
        self._cost_kw = dict(Rdiag=Rdiag, Hp=np.asarray(Ps, dtype=np.float64),
                             Qdiag_b=None if Ws_b is None else np.asarray(Ws_b, dtype=np.float64),
                             Hp_b=None if Ps_b is None else np.asarray(Ps_b, dtype=np.float64))
        self._plan_cache.clear()

    def compute_Rr_Qr(self, rho_x, rho_u, dp=True):
        """isls/base.py:55-79, returned as per-step DIAGONALS Qr[N,n], Rr[N,m] (None stays None)."""
        def ex(rho, dim):
            if rho is None:
                return None
            r = np.asarray(rho, dtype=np.float64)
            if r.ndim == 0:
                return np.full((self.N, dim), float(r))
            if r.ndim >= 2 and r.shape[-1] == r.shape[-2] == dim and not (r.ndim == 2 and r.shape[0] == self.N
                                                                           and self.N != dim):
                r = diag_of(r, "rho")
            return np.ascontiguousarray(np.broadcast_to(r, (self.N, dim)))
        return ex(rho_x, self.x_dim), ex(rho_u, self.u_dim)

    # ------------------------------------------------------------------ nominal trajectory
    @property
    def nominal_values(self):
        return self.x_nom, self.u_nom

    @nominal_values.setter
    def nominal_values(self, value):
        """(x_nom, u_nom) (isls/isls_base.py:80-85).  Only x_nom[..., 0, :] and u_nom are used: the solvers re-roll
        the model out from the initial state, exactly what the reference's rollouts do."""
        x_nom, u_nom = value
        x_nom = torch.as_tensor(np.asarray(x_nom) if not isinstance(x_nom, torch.Tensor) else x_nom)
        u_nom = torch.as_tensor(np.asarray(u_nom) if not isinstance(u_nom, torch.Tensor) else u_nom)
        self.set_initial(x_nom[..., 0, :], u_nom)

    def set_initial(self, x0, u_init):
        """x0 [n] / [B,n], u_init [N,m] / [B,N,m] (host or device)."""
        x0 = torch.as_tensor(np.asarray(x0, dtype=np.float64)) if not isinstance(x0, torch.Tensor) else x0
        u_init = (torch.as_tensor(np.asarray(u_init, dtype=np.float64)) if not isinstance(u_init, torch.Tensor)
                  else u_init)
        self._x0 = x0.reshape(-1, self.x_dim).expand(self.nb, self.x_dim)
        self._u_init = u_init.expand(self.nb, self.N, self.u_dim)

    def reset(self):
        """isls/isls_base.py:160-175."""
        self.x_nom = self.u_nom = None
        self.cost = None
        self.cost_log = None
        self._K = self._k = None
        self.status = None
        self.last = None

    @property
    def K(self):
        return self._K

    @property
    def k(self):
        return self._k

    # ------------------------------------------------------------------ plumbing
    def _dt(self):
        return float(self._model_kw.get("dt", 0.01))

    def _zs_b(self):
        zs = torch.as_tensor(self.zs)
        return zs.expand(self.nb, zs.shape[-2], self.x_dim)

    def _solver(self, L, rho_x, bx, rho_u, bu, max_outer, max_admm, want_gains=False, want_masks=False,
                obstacles=None, isls_dim=0):
        if self._model is None or self.zs is None:
            raise IslsError("set forward_model and a cost (set_quadratic_cost / cost_function) first")
        key = (L, None if rho_x is None else rho_x.tobytes(), None if bx is None else (bx[0].tobytes(), bx[1].tobytes()),
               None if rho_u is None else rho_u.tobytes(), None if bu is None else (bu[0].tobytes(), bu[1].tobytes()),
               max_outer, max_admm, want_gains, want_masks, None if obstacles is None else obstacles.key(), isls_dim)
        if key not in self._plan_cache:
            plan = S.Plan(self._model, self.N, self.x_dim, self.u_dim, self._dt(), self.Qdiag, self.seq, self.u_std,
                          L, rho_x=rho_x, lo_x=None if bx is None else bx[0], hi_x=None if bx is None else bx[1],
                          rho_u=rho_u, lo_u=None if bu is None else bu[0], hi_u=None if bu is None else bu[1],
                          cost=self._cost, obstacles=None if obstacles is None else obstacles.as_dict(),
                          isls_dim=isls_dim, device=self.device, **self._cost_kw)
            self._plan_cache.clear()
            self._plan_cache[key] = S.BatchSolver(plan, self.nb, self.device, max_outer=max_outer, max_admm=max_admm,
                                                  want_gains=want_gains, want_masks=want_masks)
        return self._plan_cache[key]

    def _publish(self, out):
        # the solver state persists like the reference's (x_nom, u_nom): a following call continues from here
        self._x0 = out.x[:, 0].clone()
        self._u_init = out.u.clone()
        # published attributes are COPIES: the solver's result buffers (`out`, also returned by the solve methods and kept
        # as self.last) are reused by the next solve on this object, the attributes a caller kept are not touched by it
        sq = (lambda t: t[0].clone()) if self.batch is None else (lambda t: t.clone())
        self.x_nom, self.u_nom = sq(out.x), sq(out.u)
        self.cost = sq(out.cost)
        self.cost_log = sq(out.cost_log)
        self.status = sq(out.status)
        if "K" in out:
            self._K, self._k = sq(out.K), sq(out.k)
        self.last = out

    # ------------------------------------------------------------------ solvers
    def solve(self, get_AB=None, get_Cs=None, is_dynamics_linear=False, is_cost_quadratic=False, method="dp",
              max_iter=100, max_line_search_iter=25, tol_fun=1e-5, tol_grad=1e-4, verbose=False, fixed_budget=False):
        """iLQR with the Riccati backward pass and closed-loop line search (isls/isls.py:54-132, method='dp')."""
        if method != "dp":
            raise NotImplementedError("device path implements method='dp' (the dense batch form is the same minimiser)")
        self._check_get_AB(get_AB)
        self._check_get_Cs(get_Cs)
        sv = self._solver(max_line_search_iter, None, None, None, None, max_iter, 1, want_gains=True)
        sv.set_inputs(self._x0, self._u_init, self._zs_b())
        out = sv.ilqr(tol_fun=tol_fun, fixed_budget=fixed_budget)
        self._publish(out)
        if verbose:
            self._report(out)
        return out

    def iterate_once_dp(self, max_line_search=15, verbose=False, **kwargs):
        """One iLQR iteration from the current nominal trajectory: backward pass, closed-loop line search, accept iff
        the cost decreases (isls/isls.py:336-374).  Returns (fp_success [B] bool, K, k)."""
        out = self.solve(method="dp", max_iter=1, max_line_search_iter=max_line_search, verbose=verbose,
                         fixed_budget=True)
        ok = out.alpha_idx[:, 0, 0] >= 0
        return (bool(ok[0]) if self.batch is None else ok), self._K, self._k

    def solve_ilqr(self, get_AB=None, max_ilqr_iter=100, max_line_search_iter=25, dp=True, verbose=False, **kw):
        """Legacy spelling used by the notebooks (Car/Iterative LQR with state constraints.ipynb cell 13).  dp=False
        asked the reference for its dense batch least-squares form (isls/isls.py:157-227): not built on the device -
        raises like solve(method='batch') instead of silently running the DP form."""
        return self.solve(get_AB, method="dp" if dp else "batch", max_iter=max_ilqr_iter,
                          max_line_search_iter=max_line_search_iter, verbose=verbose, **kw)

    def ilqr_admm(self, get_AB=None, get_Cs=None, project_x=False, project_u=False, max_iter=20,
                  max_line_search_iter=20, max_admm_iter=20, rho_x=None, rho_u=None, alpha=1, tol=1e-3, verbose=False,
                  log=False, k_max=None, max_line_search=None, threshold=None, fixed_budget=False, want_masks=False):
        """iLQR-ADMM with box state / control bounds (isls/isls.py:379-501 + isls/admm.py:6-106).
        project_x / project_u: `Bound` descriptors (isls_b200.projections).  Legacy keywords k_max,
        max_line_search, threshold map to max_iter, max_line_search_iter, tol."""
        if k_max is not None:
            max_iter = k_max
        if max_line_search is not None:
            max_line_search_iter = max_line_search
        if threshold is not None:
            tol = threshold
        self._check_get_AB(get_AB)
        self._check_get_Cs(get_Cs)
        obstacles = project_x if isinstance(project_x, ObstacleSets) else None
        for nm, pr in (("project_x", project_x), ("project_u", project_u)):
            if pr and not isinstance(pr, Bound) and not (nm == "project_x" and obstacles is not None):
                raise TypeError("%s must be an isls_b200.projections.Bound (box bounds) or, for project_x, ObstacleSets; "
                                "Python callables cannot run inside the kernels" % nm)
        Qr, Rr = self.compute_Rr_Qr(rho_x if project_x else None, rho_u if project_u else None)
        if project_x and Qr is None or project_u and Rr is None:
            raise ValueError("rho_x / rho_u is required for a projected variable")
        bx = project_x.expand(self.N, self.x_dim) if project_x and obstacles is None else None
        bu = project_u.expand(self.N, self.u_dim) if project_u else None
        sv = self._solver(max_line_search_iter, Qr, bx, Rr, bu, max_iter, max_admm_iter, want_masks=want_masks,
                          obstacles=obstacles)
        sv.set_inputs(self._x0, self._u_init, self._zs_b())
        out = sv.ilqr_admm(tol=tol, relax=float(alpha), fixed_budget=fixed_budget)
        self._publish(out)
        if verbose:
            self._report(out)
        return out.res_log if log else out

    def isls_admm(self, dim, get_AB=None, get_Cs=None, project_x=False, project_u=False, max_admm_iter=20, k_max=20,
                  max_line_search=20, rho_x=None, rho_u=None, alpha=1, threshold=1e-3, verbose=False, log=False,
                  fixed_budget=False):
        """Robust nonlinear iSLS-ADMM (isls/isls.py:503-712): ADMM on [d_u | Phi_u(:, :dim)] - robustness with respect
        to the first `dim` components of the initial state - with the row-wise SOC projection of the notebooks
        (project_u: `SetConvexSOC`, the device form of `project_u(z, u_nom)` in 3DoF robot/State bounds and robust
        control bounds.ipynb cell 25).  Returns (du [B, N m], phi_u [B, N m, dim]) like the reference."""
        self._check_get_AB(get_AB)
        self._check_get_Cs(get_Cs)
        if project_x and not isinstance(project_x, SetConvexSOCComponents):
            raise TypeError("project_x must be an isls_b200.projections.SetConvexSOCComponents; Python callables cannot "
                            "run inside the kernels")
        if project_u and not isinstance(project_u, SetConvexSOC):
            raise TypeError("project_u must be an isls_b200.projections.SetConvexSOC; Python callables cannot run "
                            "inside the kernels")
        if project_u:
            _, Rr = self.compute_Rr_Qr(None, rho_u)
            if Rr is None:
                raise ValueError("rho_u is required")
        else:                                # isls_admm without constraints (notebook cell 23): Rr = 0, z = x
            Rr = np.zeros((self.N, self.u_dim))
            project_u = None
        inf = np.full((self.N, self.u_dim), np.inf)
        Qr, bx = None, None
        if project_x:
            Qr, _ = self.compute_Rr_Qr(rho_x, None)
            if Qr is None:
                raise ValueError("rho_x is required")
            infx = np.full((self.N, self.x_dim), np.inf)
            bx = (-infx, infx)
        else:
            project_x = None
        sv = self._solver(max_line_search, Qr, bx, Rr, (-inf, inf), k_max, max_admm_iter, isls_dim=int(dim))
        sv.set_inputs(self._x0, self._u_init, self._zs_b())
        out = sv.isls_admm(project_u, tol=threshold, relax=float(alpha), fixed_budget=fixed_budget, soc_x=project_x)
        self._publish(out)
        if verbose:
            self._report(out)
        B_ = self.nb
        du = out.d_u.reshape(B_, self.N * self.u_dim)
        phi = out.phi_u.reshape(B_, self.N * self.u_dim, int(dim))
        return (du[0], phi[0]) if self.batch is None else (du, phi)

    def get_AB(self, x_nom=None, u_nom=None):
        """A, B = get_AB(x_nom, u_nom): the Jacobians of the device model along a trajectory (default: the current
        nominal values), A [N, n, n], B [N, n, m] - what the reference's callers compute with their own `get_AB`
        callable and hand to `iSLS.AB` (isls/isls_base.py:133-158)."""
        if self._model is None:
            raise IslsError("set forward_model first")
        if x_nom is None:
            if self.x_nom is None:
                raise IslsError("no nominal values yet: pass x_nom, u_nom or solve first")
            x_nom, u_nom = self.x_nom, self.u_nom
        return S.linearize(self._model, self.x_dim, self.u_dim, self._dt(), x_nom, u_nom, device=self.device)

    def controller(self, PHI_U, du):
        """K, k = controller(PHI_U, du) - the call the robust notebook makes on an iSLS object after isls_admm
        (3DoF robot/State bounds and robust control bounds.ipynb cells 23, 26; README "iSLS.controller"; the formula is
        SLS.controller, isls/sls.py:235-242, on the time-varying operators C, D): Phi_x = C + D Phi_u,
        K = Phi_u Phi_x^-1, k = (I - K D) du.

        For the Phi_u that isls_admm returns - non-zero only in its first `dim` columns (robustness with respect to the
        initial state) - this is an identity, not a solve: with E the first `dim` rows of the identity, Phi_u = phi E,
        Phi_x = C + (D phi) E, and because the first block row of D is zero (x_0 does not depend on the controls) and
        that of C^-1 = I - Z A is [I 0 ...], Woodbury gives E Phi_x^-1 = E, hence K = phi E = Phi_u and
        k = du - phi (E D du) = du.  (The reference's dense inverse reproduces exactly that: |K - Phi_u| = 1.4e-15 on the
        notebook problem, tests/golden/make_golden.py.)

        Any other causal (block lower triangular) Phi_u goes through the block back-substitution on the device
        (isls_controller_tv_f64) with C, D built from A_t, B_t: the pair set through `iSLS.AB` if there is one (the
        reference uses the C, D of its last `self.AB = ...`, isls/isls_base.py:138-158), otherwise the Jacobians of the
        device model at the current nominal values (`get_AB()`)."""
        if self.batch is not None:
            raise IslsError("controller evaluates ONE problem: construct iSLS without `batch`")
        dev = self.device
        PHI_U = torch.as_tensor(PHI_U, dtype=torch.float64).to(dev)
        du = torch.as_tensor(du, dtype=torch.float64).to(dev)
        Nm, Nn = self.N * self.u_dim, self.N * self.x_dim
        if PHI_U.shape != (Nm, Nn) or du.shape != (Nm,):
            raise ValueError("PHI_U must be [N m, N n] and du [N m]")
        if not bool((PHI_U[:, self.x_dim:] != 0).any()):
            return PHI_U.clone(), du.clone()
        if getattr(self, "A", None) is not None and getattr(self, "B", None) is not None:
            A_, B_ = self.A, self.B
        else:
            A_, B_ = self.get_AB()
        return S.controller_tv(A_, B_, PHI_U, du, device=dev)

    def _report(self, out):
        st = out.status.cpu().numpy()
        it = out.outer_iters.cpu().numpy()
        for b in range(min(self.nb, 8)):
            msg = []
            if st[b] & S.ST_CONVERGED_COST:
                msg.append("Cost change is too low, cannot improve anymore at iteration %d." % it[b])
            if st[b] & S.ST_OSCILLATING:
                msg.append("Cost is oscillating at iteration %d" % it[b])
            if st[b] & S.ST_LINESEARCH_FAIL:
                msg.append("Forward pass failed, cannot improve anymore at iteration %d." % it[b])
            if st[b] & S.ST_MAX_ITER:
                msg.append("Maximum iterations reached.")
            print("problem", b, "iLQR cost:", float(out.cost[b]), " ".join(msg))

    # ------------------------------------------------------------------ stage-level API
    def backward_pass_DP(self, Cts, cts):
        """K, k = backward_pass_DP(Cts, cts) with self.A, self.nb set (isls/isls.py:229-308)."""
        dev = self.device
        t = lambda a: torch.as_tensor(a, dtype=torch.float64).to(dev)
        A, Bm, C, c = t(self.A), t(self.B), t(Cts), t(cts)
        single = A.ndim == 3
        if single:
            A, Bm, C, c = A[None], Bm[None], C[None], c[None]
        K, k, bad = S.riccati(A, Bm, c, C)
        if bool(bad.any()):
            raise np.linalg.LinAlgError("Quu is not positive definite")       # what dposv raises, isls.py:296
        return (K[0], k[0]) if single else (K, k)

    @property
    def AB(self):
        return [self.A, self.B]

    @AB.setter
    def AB(self, value):
        self.A, self.B = value[0], value[1]

    def rollout_DP(self, K, k):
        """Closed-loop rollouts u_t = K_t (x_t - x^_t) + k_t + u^_t around the current nominal trajectory for every
        feed-forward sequence k[l] (isls/isls.py:310-334; k: [L, N, m] = k_t scaled by the line-search step sizes).
        Single-problem objects only (like the reference); returns x_log [L, N, n], u_log [L, N, m]."""
        if self.batch is not None or self.x_nom is None:
            raise IslsError("rollout_DP needs a single-problem iSLS with nominal values (after solve / iterate_once_dp)")
        dev = self.device
        t = lambda a: torch.as_tensor(a, dtype=torch.float64).to(dev)
        K, k = t(K), t(k)
        if k.ndim == 2:
            k = k[None]
        xn, un = self.x_nom, self.u_nom
        kabs = k + (un - torch.einsum("tij,tj->ti", K, xn))[None]             # absolute form u = K x + k'
        xs, us = [], []
        for l in range(k.shape[0]):
            x, u = S.mc_rollout(self._model, self.x_dim, self.u_dim, self.N, self._dt(), "dp", xn[:1], K, kabs[l],
                                device=dev)
            xs.append(x[0]); us.append(u[0])
        return torch.stack(xs), torch.stack(us)

    def rollout_batch(self, x_nom, u_nom):
        """Open-loop rollout from x_nom[0] for every control sequence in u_nom [nb,N,m] (isls/isls.py:135-154)."""
        x_nom = torch.as_tensor(np.asarray(x_nom)) if not isinstance(x_nom, torch.Tensor) else x_nom
        u_nom = torch.as_tensor(np.asarray(u_nom)) if not isinstance(u_nom, torch.Tensor) else u_nom
        nb = u_nom.shape[0]
        plan = S.Plan(self._model, self.N, self.x_dim, self.u_dim, self._dt(),
                      np.zeros((1, self.x_dim)) if self.zs is None else self.Qdiag,
                      np.zeros(self.N, dtype=np.int32) if self.zs is None else self.seq,
                      0.0 if self.zs is None else self.u_std, 1, device=self.device)
        sv = S.BatchSolver(plan, nb, self.device, logs=False)
        zs = torch.zeros(nb, plan.n_via, self.x_dim, dtype=torch.float64)
        xn = x_nom.reshape(-1, self.N, self.x_dim)[:1].expand(nb, self.N, self.x_dim)
        _, _, xb, ub = sv.linesearch(xn, u_nom, torch.zeros_like(u_nom), zs)
        return xb, ub


from .sls import _mc_api      # noqa: E402
iSLS.get_trajectory_batch, iSLS.get_trajectory_dp, iSLS.get_trajectory_sls = _mc_api(True)
