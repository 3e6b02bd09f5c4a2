"""SLS - device mirror of the LQT-ADMM part of the reference's `isls.SLS` (isls/sls.py, isls/sls_base.py)."""
import numpy as np
import torch

from . import solver as S
from ._lib import IslsError
from .projections import Bound
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

    @property
    def AB(self):
        return [self.A, self.B]

    @AB.setter
    def AB(self, value):
        """Linear dynamics (isls/base.py:98-119).  The device path registers the double integrator
        (isls/utils.py:266-276): A, B must be get_double_integrator_AB(u_dim, 2, dt) for some dt."""
        A, Bm = np.asarray(value[0], dtype=np.float64), np.asarray(value[1], dtype=np.float64)
        d = self.u_dim
        if A.shape != (2 * d, 2 * d) or Bm.shape != (2 * d, d) or self.x_dim != 2 * d:
            raise NotImplementedError("device LQT path supports the double integrator (x_dim = 2 u_dim)")
        dt = float(A[0, d])
        A2, B2 = get_double_integrator_AB(d, 2, dt)
        if not (np.array_equal(A, A2) and np.array_equal(Bm, B2)):
            raise NotImplementedError("A, B are not a double integrator; no device model registered for them")
        self.A, self.B, self._dt = A, Bm, dt

    def set_quadratic_cost(self, zs, Qs, seq, u_std):
        """isls/base.py:81-89."""
        self.zs = np.asarray(zs, dtype=np.float64)
        Qs = np.asarray(Qs, dtype=np.float64)
        self.Qdiag = diag_of(Qs, "Qs") if Qs.ndim == 3 else Qs
        self.seq = np.asarray(seq, dtype=np.int32)
        self.u_std = float(u_std)

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

    def ADMM_LQT_DP(self, x0, project_x=False, project_u=False, max_iter=2000, rho_x=None, rho_u=None, alpha=1.0,
                    tol=1e-3, verbose=False, log=False, fixed_budget=False, want_masks=False):
        """LQT-ADMM with dynamic programming (isls/sls.py:298-317): Riccati pass once, then ff-pass + rollout +
        projection/dual update per iteration, all inside one kernel.  Returns (x, u, K, k[, logs])."""
        if self._dt is None or self.zs is None:
            raise IslsError("set AB and set_quadratic_cost first")
        for nm, pr in (("project_x", project_x), ("project_u", project_u)):
            if pr and not isinstance(pr, Bound):
                raise TypeError("%s must be an isls_b200.projections.Bound" % nm)
        bx = project_x.expand(self.N, self.x_dim) if project_x else None
        bu = project_u.expand(self.N, self.u_dim) if project_u else None
        plan = S.Plan("double_integrator", self.N, self.x_dim, self.u_dim, self._dt, self.Qdiag, self.seq, self.u_std,
                      1, rho_x=self._rho(rho_x, self.x_dim) if project_x else None,
                      lo_x=None if bx is None else bx[0], hi_x=None if bx is None else bx[1],
                      rho_u=self._rho(rho_u, self.u_dim) if project_u else None,
                      lo_u=None if bu is None else bu[0], hi_u=None if bu is None else bu[1])
        sv = S.BatchSolver(plan, self.nb, self.device, max_outer=1, max_admm=max_iter, logs=True, want_gains=True,
                           want_masks=want_masks)
        x0 = torch.as_tensor(np.asarray(x0, dtype=np.float64)) if not isinstance(x0, torch.Tensor) else x0
        zs = torch.as_tensor(self.zs)
        sv.set_inputs(x0.reshape(-1, self.x_dim).expand(self.nb, self.x_dim),
                      torch.zeros(self.N, self.u_dim, dtype=torch.float64),
                      zs.expand(self.nb, zs.shape[-2], self.x_dim))
        out = sv.lqt_admm_dp(tol=tol, relax=float(alpha), fixed_budget=fixed_budget)
        self.last = out
        sq = (lambda t: t[0]) if self.batch is None else (lambda t: t)
        ret = (sq(out.x).reshape(*out.x.shape[:-2], -1) if self.batch is not None else out.x[0].reshape(-1),
               sq(out.u).reshape(*out.u.shape[:-2], -1) if self.batch is not None else out.u[0].reshape(-1),
               sq(out.K), sq(out.k))
        if log:
            ret += (sq(out.res_log[:, 0]),)
        return ret
