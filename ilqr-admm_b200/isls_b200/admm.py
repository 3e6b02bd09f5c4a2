"""ADMM - generic scaled-dual ADMM driver of the reference (isls/admm.py:6-106) on CUDA tensors.

`f_argmin(reg_x, reg_u) -> (x_x, x_u, ...)` is a user callable producing flat CUDA float64 tensors [len] (or
[B, len]); the z-projection + dual update + residual sums run in the isls_admm_project_dual_f64 kernel with
`Bound` projections.  The iLQR-ADMM / LQT-ADMM solvers do not go through this host loop (their loop is on the
device); it exists for API parity and for the stage-level tests.
"""
import torch

from . import solver as S
from .projections import Bound


def ADMM(shape_x, shape_u, f_argmin, project_x=False, project_u=False, z_x_init=None, z_u_init=None,
         lmb_x_init=None, lmb_u_init=None, Qr=None, Rr=None, max_iter=20, alpha=1.0, tol=1e-3, verbose=False,
         return_lmb=False, log=False, device="cuda:0"):
    f64 = dict(dtype=torch.float64, device=device)
    logs = []

    def init(v, shape, on):
        if not on:
            return None
        return (torch.zeros(shape, **f64) if v is None else torch.as_tensor(v, **f64).clone()).reshape(1, -1)
    for nm, pr in (("project_x", project_x), ("project_u", project_u)):
        if pr and not isinstance(pr, Bound):
            raise TypeError("%s must be an isls_b200.projections.Bound" % nm)
    z_x, z_u = init(z_x_init, shape_x, project_x), init(z_u_init, shape_u, project_u)
    l_x, l_u = init(lmb_x_init, shape_x, project_x), init(lmb_u_init, shape_u, project_u)
    bx = [torch.as_tensor(a.reshape(-1), **f64) for a in project_x.expand(1, shape_x)] if project_x else None
    bu = [torch.as_tensor(a.reshape(-1), **f64) for a in project_u.expand(1, shape_u)] if project_u else None
    prim = dual = 1e6
    ret = None
    for j in range(max_iter):
        reg_x = (z_x - l_x)[0] if project_x else None                         # admm.py:32-33
        reg_u = (z_u - l_u)[0] if project_u else None
        ret = f_argmin(reg_x, reg_u)
        x_x, x_u = ret[0], ret[1]
        pprim, pdual = prim, dual
        sq = []                                                               # squared residual sums, still on the device
        if project_x:
            sq += list(S.admm_project_dual(x_x.reshape(1, -1).contiguous(), z_x, l_x, bx[0], bx[1], alpha)[:2])
        if project_u:
            sq += list(S.admm_project_dual(x_u.reshape(1, -1).contiguous(), z_u, l_u, bu[0], bu[1], alpha)[:2])
        prim = dual = 0.0
        if sq:                                                                # ONE device -> host read per iteration
            r = torch.cat(sq).sqrt().tolist()
            prim, dual = sum(r[0::2]), sum(r[1::2])                           # admm.py:62-69
        logs.append((prim, dual))
        if prim < tol and dual < tol:                                         # admm.py:72
            if verbose:
                print("ADMM converged at iteration ", j, "!")
            break
        pch = abs(pprim - prim) / (pprim + 1e-30)                             # admm.py:78-79
        dch = abs(pdual - dual) / (pdual + 1e-30)
        if pch < tol and dch < tol:
            if verbose:
                print("ADMM can't improve anymore at iteration ", j, "!")
            break
    out = tuple(ret)
    if return_lmb:
        sq = lambda t: None if t is None else t[0]
        out += (sq(l_x), sq(l_u), sq(z_x), sq(z_u))
    if log:
        out += (logs,)
    return out
