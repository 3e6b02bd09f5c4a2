"""SLS - device mirror of the LQT-ADMM part of the reference's `isls.SLS` (isls/sls.py, isls/sls_base.py)."""
import numpy as np
import torch

from . import solver as S
from ._lib import IslsError
import ctypes as C

from . import _lib
from .projections import Bound, ObstacleSets, SetConvexSOC, SetConvexSOCRows
from .utils import diag_of, get_double_integrator_AB


class SLS:
    def __init__(self, x_dim, u_dim, N, batch=None, device="cuda:0"):
        """SLS(x_dim, u_dim, N) (isls/sls.py:9-38)."""
        self.x_dim, self.u_dim, self.N = int(x_dim), int(u_dim), int(N)
        self.batch = batch
        self.nb = 1 if batch is None else int(batch)
        self.device = device
        self.A = self.B = None
        self._dt = None
        self.zs = None
        self.last = None
        self._sls_plan = None

    @property
    def AB(self):
        return [self.A, self.B]

    @AB.setter
    def AB(self, value):
        """Linear dynamics (isls/base.py:98-119): any constant pair A [x_dim, x_dim], B [x_dim, u_dim].  A double
        integrator (isls/utils.py:266-276) runs on the registered sparse model, anything else on the dense "lti" model
        ((x_dim, u_dim) in (2,1), (4,2), (6,3))."""
        A, Bm = np.asarray(value[0], dtype=np.float64), np.asarray(value[1], dtype=np.float64)
        d = self.u_dim
        if A.shape != (self.x_dim, self.x_dim) or Bm.shape != (self.x_dim, d):
            raise ValueError("device SLS path needs constant A [x_dim, x_dim], B [x_dim, u_dim]")
        self.A, self.B, self._dt = A, Bm, None
        self._sls_plan = None
        if self.x_dim == 2 * d:                      # registered device model for the LQT-ADMM (DP) kernels
            dt = float(A[0, d])
            A2, B2 = get_double_integrator_AB(d, 2, dt)
            if np.array_equal(A, A2) and np.array_equal(Bm, B2):
                self._dt = dt

    def set_quadratic_cost(self, zs, Qs, seq, u_std):
        """isls/base.py:81-89."""
        self.zs = np.asarray(zs, dtype=np.float64)
        Qs = np.asarray(Qs, dtype=np.float64)
        self.Qdiag = diag_of(Qs, "Qs") if Qs.ndim == 3 else Qs
        self.seq = np.asarray(seq, dtype=np.int32)
        self.u_std = float(u_std)
        self._sls_plan = None

    set_cost_variables = set_quadratic_cost

    def _rho(self, rho, dim):
        if rho is None:
            return None
        r = np.asarray(rho, dtype=np.float64)
        if r.ndim == 0:
            return np.full((self.N, dim), float(r))
        if r.ndim >= 2 and r.shape[-1] == r.shape[-2] == dim:
            r = diag_of(r, "rho")
        return np.ascontiguousarray(np.broadcast_to(r, (self.N, dim)))

    def _lqt(self, x0, project_x, project_u, max_iter, rho_x, rho_u, alpha, tol, fixed_budget, want_masks,
             last_stage_dp=True, z_init=None, want_Qs=False):
        obst = project_x if isinstance(project_x, ObstacleSets) else None
        bx = project_x.expand(self.N, self.x_dim) if project_x and obst is None else None
        bu = project_u.expand(self.N, self.u_dim) if project_u else None
        lti = self._dt is None
        plan = S.Plan("lti" if lti else "double_integrator", self.N, self.x_dim, self.u_dim, 0.0 if lti else self._dt,
                      self.Qdiag, self.seq, self.u_std, 1, lti_AB=(self.A, self.B) if lti else None,
                      rho_x=self._rho(rho_x, self.x_dim) if project_x else None,
                      lo_x=None if bx is None else bx[0], hi_x=None if bx is None else bx[1],
                      rho_u=self._rho(rho_u, self.u_dim) if project_u else None,
                      lo_u=None if bu is None else bu[0], hi_u=None if bu is None else bu[1],
                      obstacles=None if obst is None else obst.as_dict(), device=self.device)
        sv = S.BatchSolver(plan, self.nb, self.device, max_outer=1, max_admm=max_iter, logs=True, want_gains=True,
                           want_masks=want_masks, want_Qs=want_Qs)
        x0 = torch.as_tensor(np.asarray(x0, dtype=np.float64)) if not isinstance(x0, torch.Tensor) else x0
        zs = torch.as_tensor(self.zs)
        sv.set_inputs(x0.reshape(-1, self.x_dim).expand(self.nb, self.x_dim),
                      torch.zeros(self.N, self.u_dim, dtype=torch.float64),
                      zs.expand(self.nb, zs.shape[-2], self.x_dim))
        out = sv.lqt_admm_dp(tol=tol, relax=float(alpha), fixed_budget=fixed_budget, last_stage_dp=last_stage_dp,
                             z_x_init=None if z_init is None else z_init[0], z_u_init=None if z_init is None else z_init[1])
        self.last = out
        return out

    def _check_lqt(self):
        if self.A is None or self.zs is None:
            raise IslsError("set AB and set_quadratic_cost first")
        if self._dt is None and (self.x_dim, self.u_dim) not in ((2, 1), (4, 2), (6, 3)):
            raise NotImplementedError("the dense LTI model of the LQT (DP) kernels is compiled for (x_dim, u_dim) in "
                                      "(2,1), (4,2), (6,3)")

    def _regularised(self, Qr, Rr, ur, xr, x0, return_Qs):
        """One K-pass + one feed-forward pass of the regularised LQT problem (isls/sls.py:85-202 with Qr, Rr, xr, ur):
        on the device this is the first ADMM iteration of the LQT kernel started from z = (xr, ur), lambda = 0 with
        penalties diag(Qr), diag(Rr) and identity projections.  Qr [N,n,n] / Rr N x [m,m] must be diagonal."""
        self._check_lqt()
        N, n, m = self.N, self.x_dim, self.u_dim
        inf = Bound(-np.inf, np.inf)
        rho_x = rho_u = None
        zx = zu = None
        f64 = dict(dtype=torch.float64, device=self.device)
        if Qr is not None:
            if xr is None:
                raise ValueError("Qr needs xr (isls/sls.py:108)")
            rho_x = diag_of(np.asarray(Qr, dtype=np.float64), "Qr")
            zx = torch.as_tensor(np.asarray(xr, dtype=np.float64).reshape(-1, N, n), **f64).expand(self.nb, N, n)
        if Rr is not None:
            if ur is None:
                raise ValueError("Rr needs ur (isls/sls.py:114)")
            rho_u = diag_of(np.stack([np.asarray(r, dtype=np.float64) for r in Rr]), "Rr")
            zu = torch.as_tensor(np.asarray(ur, dtype=np.float64).reshape(-1, N, m), **f64).expand(self.nb, N, m)
        return self._lqt(np.zeros(n) if x0 is None else x0, inf if Qr is not None else False,
                         inf if Rr is not None else False, 1, rho_x, rho_u, 1.0, 0.0, True, False,
                         z_init=(zx, zu) if (zx is not None or zu is not None) else None, want_Qs=return_Qs)

    def solve_dp(self, Qr=None, Rr=None, ur=None, xr=None, return_Qs=False, x0=None):
        """K, k of the LQT problem by the Riccati recursion (isls/sls.py:85-166), optionally regularised for ADMM
        (Cxx += 2Qr, cx -= 2Qr xr, Cuu += 2Rr, cu -= 2Rr ur; diagonal Qr / Rr) and with the logs Quu, Quu_inv, Qux
        (return_Qs) that solve_dp_ff takes."""
        out = self._regularised(Qr, Rr, ur, xr, x0, return_Qs)
        sq = (lambda t: t[0].clone()) if self.batch is None else (lambda t: t.clone())
        if return_Qs:
            return sq(out.K), sq(out.k), sq(out.Quu), sq(out.Quu_inv), sq(out.Qux)
        return sq(out.K), sq(out.k)

    def solve(self, x0=None, method="dp", verbose=False):
        """x, u of the LQT problem from x0 (isls/sls.py:40-60; method 'dp': Riccati gains + closed-loop rollout,
        sls_base.py:76-89)."""
        if method != "dp":
            raise NotImplementedError("device path implements method='dp'")
        self._check_lqt()
        out = self._lqt(np.zeros(self.x_dim) if x0 is None else x0, False, False, 1, None, None, 1.0, 0.0, True, False)
        if self.batch is None:
            return out.x[0].reshape(-1), out.u[0].reshape(-1)
        return out.x.reshape(self.nb, -1), out.u.reshape(self.nb, -1)

    def compute_cost(self, x, u=None):
        """sum (x-xd)'Q(x-xd) + u'Ru, batched over leading axes (isls/sls_base.py:25-44).  Convenience only
        (elementwise torch ops on the caller's tensors; the solvers evaluate costs inside their kernels)."""
        x = torch.as_tensor(x, dtype=torch.float64)
        dev = x.device
        Qt = torch.as_tensor(np.ascontiguousarray(self.Qdiag[self.seq]).reshape(-1), device=dev)
        zs = np.asarray(self.zs, dtype=np.float64)
        xd = zs[..., self.seq, :].reshape(*zs.shape[:-2], -1)
        xd = torch.as_tensor(np.ascontiguousarray(xd), device=dev)
        dx = x.reshape(*x.shape[:-2], -1) - xd if x.shape[-1] == self.x_dim and x.ndim >= 2 else x - xd
        c = (dx * dx * Qt).sum(-1)
        if u is not None:
            u = torch.as_tensor(u, dtype=torch.float64, device=dev)
            uf = u.reshape(*u.shape[:-2], -1) if u.shape[-1] == self.u_dim and u.ndim >= 2 else u
            c = c + self.u_std * (uf * uf).sum(-1)
        return c

    def ADMM_LQT_DP(self, x0, project_x=False, project_u=False, max_iter=2000, rho_x=None, rho_u=None, alpha=1.0,
                    tol=1e-3, verbose=False, log=False, fixed_budget=False, want_masks=False):
        """LQT-ADMM with dynamic programming (isls/sls.py:298-317): Riccati pass once, then ff-pass + rollout +
        projection/dual update per iteration, all inside one kernel.  Returns (x, u, K, k[, logs])."""
        self._check_lqt()
        for nm, pr in (("project_x", project_x), ("project_u", project_u)):
            if pr and not isinstance(pr, Bound) and not (nm == "project_x" and isinstance(pr, ObstacleSets)):
                raise TypeError("%s must be an isls_b200.projections.Bound (or ObstacleSets for project_x)" % nm)
        out = self._lqt(x0, project_x, project_u, max_iter, rho_x, rho_u, alpha, tol, fixed_budget, want_masks)
        sq = (lambda t: t[0]) if self.batch is None else (lambda t: t)
        ret = (sq(out.x).reshape(*out.x.shape[:-2], -1) if self.batch is not None else out.x[0].reshape(-1),
               sq(out.u).reshape(*out.u.shape[:-2], -1) if self.batch is not None else out.u[0].reshape(-1),
               sq(out.K), sq(out.k))
        if log:
            ret += (sq(out.res_log[:, 0]),)
        return ret

    def ADMM_LQT_Batch(self, x0, project_x=False, project_u=False, max_iter=20, rho_x=None, rho_u=None, alpha=1.0,
                       tol=1e-3, verbose=False, log=False, fixed_budget=False, want_masks=False):
        """LQT-ADMM in "batch" form (isls/sls.py:250-294).  The reference's dense least-squares argmin is the same LQ
        minimiser the Riccati recursion computes, except that its last control is solved for (u_{N-1} = (R + Rr)^-1 Rr
        reg_u) and ADMM starts from the unconstrained solution (sls.py:266-268); both are reproduced here on top of
        the LQT kernel.  Returns (x, u[, logs])."""
        self._check_lqt()
        for nm, pr in (("project_x", project_x), ("project_u", project_u)):
            if pr and not isinstance(pr, Bound) and not (nm == "project_x" and isinstance(pr, ObstacleSets)):
                raise TypeError("%s must be an isls_b200.projections.Bound (or ObstacleSets for project_x)" % nm)
        unc = self._lqt(x0, False, False, 1, None, None, 1.0, 0.0, True, False)          # z_x_init, z_u_init
        z_init = (unc.x.clone(), unc.u.clone())
        out = self._lqt(x0, project_x, project_u, max_iter, rho_x, rho_u, alpha, tol, fixed_budget, want_masks,
                        last_stage_dp=False, z_init=z_init)
        sq = (lambda t: t[0]) if self.batch is None else (lambda t: t)
        ret = (sq(out.x).reshape(*out.x.shape[:-2], -1) if self.batch is not None else out.x[0].reshape(-1),
               sq(out.u).reshape(*out.u.shape[:-2], -1) if self.batch is not None else out.u[0].reshape(-1))
        if log:
            ret += (sq(out.res_log[:, 0]),)
        return ret

    def solve_dp_ff(self, K=None, Quu=None, Qux=None, Quu_inv=None, Qr=None, Rr=None, ur=None, xr=None, x0=None):
        """Feed-forward gains k of the (regularised) LQT problem for new xr, ur (isls/sls.py:168-202).  The reference
        re-uses the logs K, Quu, Qux, Quu_inv of a previous solve_dp with the same Qr, Rr; on the device that K-pass
        is part of the same launch sequence (it depends on A, B, Q, Qr, Rr only), so the log arguments are accepted for
        signature compatibility and not read."""
        out = self._regularised(Qr, Rr, ur, xr, x0, False)
        return out.k[0].clone() if self.batch is None else out.k.clone()

    # ------------------------------------------------------------------ SLS (system level synthesis) path
    def _plan(self):
        """Shared operators Sw, Su, L^-1, PHI_U on the device (isls_sls_plan_create)."""
        if self.A is None or self.zs is None:
            raise IslsError("set AB and set_quadratic_cost first")
        if self._sls_plan is None:
            L_ = _lib.lib()
            Qt = np.ascontiguousarray(self.Qdiag[self.seq])                      # [N, n] per-step diagonal of Q
            A = np.ascontiguousarray(self.A)
            Bm = np.ascontiguousarray(self.B)
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                rc = L_.isls_sls_plan_create(self.x_dim, self.u_dim, self.N, A.ctypes.data, Bm.ctypes.data,
                                             Qt.ctypes.data, self.u_std, C.byref(h),
                                             C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            _lib.check(rc, "isls_sls_plan_create")
            self._sls_plan = _SlsPlan(h, self)
        return self._sls_plan

    def _xd(self):
        """xd = find_mus(zs, seq) (isls/utils.py:95-99) per problem: [B, N n] on the device."""
        zs = np.asarray(self.zs, dtype=np.float64)
        if zs.ndim == 2:
            zs = zs[None]
        xd = zs[:, self.seq].reshape(zs.shape[0], -1)
        xd = np.ascontiguousarray(np.broadcast_to(xd, (self.nb, self.N * self.x_dim)))
        return torch.from_numpy(xd.copy()).to(self.device)

    @property
    def Sw(self):
        return self._plan().Sw

    @property
    def Su(self):
        return self._plan().Su

    def solve_sls(self, verbose=False):
        """PHI_U, du = solve_sls() (isls/sls.py:205-233).  PHI_U [N m, N n] is shared by the batch; du [B, N m]."""
        pl = self._plan()
        xd = self._xd()
        du = torch.empty(self.nb, self.N * self.u_dim, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().isls_sls_solve_f64(pl.handle, self.nb, C.c_void_p(xd.data_ptr()),
                                               C.c_void_p(du.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc, "isls_sls_solve_f64")
        return pl.PHI_U, (du[0] if self.batch is None else du)

    def ADMM_SLS(self, project_x=False, project_u=False, max_iter=5000, rho_x=0., rho_u=0., alpha=1., tol=1e-3,
                 verbose=False, log=False, fixed_budget=False):
        """SLS-ADMM with robust (chance-constrained) control bounds w.r.t. the initial position
        (isls/sls.py:319-454).  project_u: `SetConvexSOC`; returns (du, phi_u[, logs])."""
        if project_x and not isinstance(project_x, SetConvexSOCRows):
            raise TypeError("project_x must be an isls_b200.projections.SetConvexSOCRows")
        if not isinstance(project_u, SetConvexSOC):
            raise TypeError("project_u must be an isls_b200.projections.SetConvexSOC")
        pl = self._plan()
        dev = self.device
        B_, Nm, Nn, c = self.nb, self.N * self.u_dim, self.N * self.x_dim, self.x_dim // 2 + 1
        As, bs = project_u.As, project_u.bs
        if As.shape[1:] != (c + 1, c):
            raise ValueError("cone matrices must be [%d, %d] (1 + x_dim/2 columns)" % (c + 1, c))
        xd = self._xd()
        f64 = dict(dtype=torch.float64, device=dev)
        du = torch.empty(B_, Nm, **f64)
        phic = torch.empty(B_, Nm, c - 1, **f64)
        logs = torch.full((B_, int(max_iter), 2), float("nan"), **f64)
        iters = torch.empty(B_, dtype=torch.int32, device=dev)
        exits = torch.empty(B_, dtype=torch.int32, device=dev)
        inner = torch.empty(B_, dtype=torch.int64, device=dev)
        o = _lib.SlsAdmmOpts(max_iter=int(max_iter), rho_u=float(rho_u), alpha=float(alpha), tol=float(tol),
                             fixed_budget=int(fixed_budget), n_cones=As.shape[0], cone_rows=c + 1,
                             As=As.ctypes.data, bs=bs.ctypes.data, inner_rho=project_u.rho,
                             inner_max_iter=project_u.max_iter, inner_threshold=project_u.threshold)
        if project_x:
            rows = np.array([r % Nn for r in project_x.rows], dtype=np.int32)
            xbs = project_x.bs
            if xbs.shape[1:] != bs.shape:
                raise ValueError("state-row cone offsets must be [rows, %d, %d]" % bs.shape)
            rx = np.asarray(rho_x, dtype=np.float64)
            if rx.ndim == 3:                                   # [N, n, n] like the reference
                rx = np.stack([np.diag(q) for q in rx]).reshape(-1)
            rx = np.ascontiguousarray(np.broadcast_to(rx.reshape(-1), (Nn,)))
            off = np.ones(Nn, dtype=bool)
            off[rows] = False
            if np.any(rx[off] != 0.0):
                raise ValueError("rho_x must be zero on the rows project_x leaves alone")
            rxr = np.ascontiguousarray(rx[rows])
            o.n_x_rows, o.x_row_idx, o.x_bs, o.rho_x_rows = len(rows), rows.ctypes.data, xbs.ctypes.data, rxr.ctypes.data
            self._keep_x = (rows, xbs, rxr)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rc = _lib.lib().isls_sls_admm_f64(pl.handle, C.byref(o), B_, p(xd), p(du), p(phic), p(logs), p(iters),
                                              p(exits), p(inner),
                                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "isls_sls_admm_f64")
        self.last = S.Result(du=du, phi_cols=phic, logs=logs, iters=iters, exit_code=exits, inner_total=inner)
        phi_u = torch.cat([phic, pl.PHI_U[None, :, c - 1:].expand(B_, Nm, Nn - (c - 1))], dim=-1)   # sls.py:450
        phi_u._isls_first_cols = c - 1      # hint for controller(): only these columns differ from the shared PHI_U
        if verbose:
            it, ex = iters.cpu().numpy(), exits.cpu().numpy()
            for b in range(min(B_, 8)):
                msg = {1: "ADMM converged at iteration %d !", 2: "ADMM can't improve anymore at iteration %d !",
                       3: "ADMM: Max iteration reached. (%d)"}[int(ex[b])] % (it[b] - 1)
                print("problem", b, msg, "residual", logs[b, it[b] - 1].tolist())
        sq = (lambda t: t[0]) if self.batch is None else (lambda t: t)
        ret = (sq(du), sq(phi_u))
        if log:
            ret += (sq(logs),)
        return ret

    def controller(self, PHI_U, du):
        """K, k = controller(PHI_U, du) (isls/sls.py:235-242); PHI_U [N m, N n] or [B, N m, N n], du [N m] / [B, N m]."""
        pl = self._plan()
        dev = self.device
        Nm, Nn = self.N * self.u_dim, self.N * self.x_dim
        PHI_U = torch.as_tensor(PHI_U, dtype=torch.float64).to(dev)
        du = torch.as_tensor(du, dtype=torch.float64).to(dev)
        single = PHI_U.ndim == 2 and du.ndim == 1
        du = du.reshape(-1, Nm).contiguous()
        B_ = du.shape[0]
        PHI = PHI_U.reshape(-1, Nm, Nn).expand(B_, Nm, Nn)
        f64 = dict(dtype=torch.float64, device=dev)
        K = torch.empty(B_, Nm, Nn, **f64)
        k = torch.empty(B_, Nm, **f64)
        # A PHI_U that ADMM_SLS returned differs from the plan's shared PHI_U in its first c - 1 columns only; then the
        # other columns of PHI_X and K are shared too and the library computes them once (k_sls_ctrl_shared).  The hint is
        # verified on the device, so a tensor the caller modified takes the general per-problem path.
        nf = getattr(PHI_U, "_isls_first_cols", None)
        if nf is not None and PHI.shape[-1] == Nn and torch.equal(PHI[:, :, nf:], pl.PHI_U[None, :, nf:].expand(B_, Nm, Nn - nf)):
            cols, ncols = PHI[:, :, :nf].contiguous(), nf
            import os
            ws = torch.empty(B_ * Nn * Nn if os.environ.get("ISLS_SLS_CTRL_DENSE") else 1, **f64)   # test switch
        else:
            cols, ncols = PHI.contiguous(), Nn
            ws = torch.empty(B_ * Nn * Nn, **f64)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rc = _lib.lib().isls_sls_controller_f64(pl.handle, B_, ncols, p(cols), p(du), p(ws), ws.numel() * 8, p(K),
                                                    p(k), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "isls_sls_controller_f64")
        return (K[0], k[0]) if single else (K, k)


def _replan_api():
    def initialize_replanning_procedure(self, K):
        """isls/sls.py:244-245.  The replan matrix (I - K Su)(Su'Q Su + R)^-1 Su'Q is never formed: the gains are kept
        and replan_feedforward applies it as four matrix-vector products per problem."""
        Nm, Nn = self.N * self.u_dim, self.N * self.x_dim
        self._replan_K = torch.as_tensor(K, dtype=torch.float64).to(self.device).reshape(-1, Nm, Nn).contiguous()

    def replan_feedforward(self, k, xd):
        """k + replan_matrix @ (xd - self.xd) (isls/sls.py:247-248); k [N m] / [B, N m], xd [N n] / [B, N n]."""
        pl = self._plan()
        dev = self.device
        Nm, Nn = self.N * self.u_dim, self.N * self.x_dim
        k = torch.as_tensor(k, dtype=torch.float64).to(dev)
        single = k.ndim == 1
        k = k.reshape(-1, Nm).contiguous()
        B_ = k.shape[0]
        xd = torch.as_tensor(xd, dtype=torch.float64).to(dev).reshape(-1, Nn).expand(B_, Nn).contiguous()
        xo = self._xd().to(dev).reshape(-1, Nn).expand(B_, Nn).contiguous()
        K = self._replan_K.expand(B_, Nm, Nn).contiguous()
        out = torch.empty_like(k)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rc = _lib.lib().isls_sls_replan_f64(pl.handle, B_, p(K), p(k), p(xd), p(xo), p(out),
                                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "isls_sls_replan_f64")
        return out[0] if single else out
    return initialize_replanning_procedure, replan_feedforward


def _mc_api(cls_is_nonlinear):
    """get_trajectory_batch / dp / sls (isls/sls_base.py:62-105, isls/isls_base.py:28-71) on the device."""
    def _args(self):
        if cls_is_nonlinear:
            return self._model, self.x_dim, self.u_dim, self.N, self._dt()
        if self._dt is None:
            raise NotImplementedError("device rollouts need the registered double-integrator model")
        return "double_integrator", self.x_dim, self.u_dim, self.N, self._dt

    def _ret(x0, xs, us):
        single = (np.ndim(x0) == 1) if not isinstance(x0, torch.Tensor) else (x0.ndim == 1)
        return (xs[0], us[0]) if single else (xs, us)

    def get_trajectory_batch(self, x0, us, noise_scale=0, seed=0):
        xs, uo = S.mc_rollout(*_args(self), "batch", x0, None, us, noise_scale=noise_scale, seed=seed, device=self.device)
        return _ret(x0, xs, uo)

    def get_trajectory_dp(self, x0, K, k, noise_scale=0, seed=0):
        xs, uo = S.mc_rollout(*_args(self), "dp", x0, K, k, noise_scale=noise_scale, seed=seed, device=self.device)
        return _ret(x0, xs, uo)

    def get_trajectory_sls(self, x0, K, k, noise_scale=0, seed=0):
        xn = un = None
        if cls_is_nonlinear:                       # iSLSBase feeds back on (x - x_nom) and adds u_nom
            xn, un = self.x_nom, self.u_nom
            if xn is not None and xn.ndim == 3:
                raise NotImplementedError("get_trajectory_sls evaluates ONE controller: construct iSLS without `batch`")
        xs, uo = S.mc_rollout(*_args(self), "sls", x0, K, k, x_nom=xn, u_nom=un, noise_scale=noise_scale, seed=seed,
                              device=self.device)
        return _ret(x0, xs, uo)
    return get_trajectory_batch, get_trajectory_dp, get_trajectory_sls


class _SlsPlan:
    """Owner of an isls_sls_plan handle + caller-owned copies of its shared device operators."""

    def __init__(self, handle, owner):
        self.handle = handle
        dev = owner.device
        Nn, Nm = owner.N * owner.x_dim, owner.N * owner.u_dim
        f64 = dict(dtype=torch.float64, device=dev)
        self.Sw, self.Su, self.PHI_U = torch.empty(Nn, Nn, **f64), torch.empty(Nn, Nm, **f64), torch.empty(Nm, Nn, **f64)
        with torch.cuda.device(dev):
            rc = _lib.lib().isls_sls_operators(handle, C.c_void_p(self.Sw.data_ptr()), C.c_void_p(self.Su.data_ptr()),
                                               C.c_void_p(self.PHI_U.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "isls_sls_operators")

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().isls_sls_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


SLS.get_trajectory_batch, SLS.get_trajectory_dp, SLS.get_trajectory_sls = _mc_api(False)
SLS.initialize_replanning_procedure, SLS.replan_feedforward = _replan_api()
