"""Projections of the hot path (isls/projections.py).  On the device the ADMM loop supports box bounds
(`project_bound`, isls/projections.py:7-11); they are described by `Bound` objects so the kernels can apply them
element-wise (a Python callable cannot run inside a kernel, and there is no CPU fallback)."""
import numpy as np
import torch


def project_bound(x, l, u):
    """np.clip(x, l, u) (isls/projections.py:7-11) on torch tensors or numpy arrays."""
    if isinstance(x, torch.Tensor):
        l = torch.as_tensor(l, dtype=x.dtype, device=x.device)
        u = torch.as_tensor(u, dtype=x.dtype, device=x.device)
        return torch.minimum(torch.maximum(x, l), u)
    return np.clip(x, l, u)


class Bound:
    """Box constraint lo <= v <= hi on a flattened [N*dim] trajectory (the closures the reference notebooks pass
    as project_x / project_u, e.g. notebooks/Car/Iterative LQR with control constraints.ipynb cell 18).

    lo, hi: scalars, [dim], [N, dim] or flat [N*dim]; +-inf marks an unconstrained element."""

    def __init__(self, lo=-np.inf, hi=np.inf):
        self.lo = np.asarray(lo, dtype=np.float64)
        self.hi = np.asarray(hi, dtype=np.float64)

    def expand(self, N, dim):
        def ex(a):
            if a.ndim == 1 and a.size == N * dim and a.size != dim:
                a = a.reshape(N, dim)
            return np.ascontiguousarray(np.broadcast_to(a, (N, dim)))
        return ex(self.lo), ex(self.hi)

    def __call__(self, z):
        return project_bound(z, self.lo.reshape(-1) if self.lo.ndim > 1 else self.lo,
                             self.hi.reshape(-1) if self.hi.ndim > 1 else self.hi)


def bound(lo=-np.inf, hi=np.inf):
    return Bound(lo, hi)


class SetConvexSOC:
    """Row-wise projection onto the intersection of second-order cones {A_i x + b_i in SOC}: the device counterpart of
    `lambda y: project_set_convex(y, As, bs, projections=[project_soc_unit]*P, rho=.., max_iter=.., threshold=..)`
    (isls/projections.py:289-374 with 118-162), e.g. the chance-constrained control bounds of
    notebooks/Double integrator/LQR and SLS with control bounds.ipynb cell 15.  As: list of [c+1, c], bs: list of
    [c+1]."""

    def __init__(self, As, bs, rho=1.0, max_iter=200, threshold=1e-4):
        self.As = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in As]))
        self.bs = np.ascontiguousarray(np.stack([np.asarray(b, dtype=np.float64) for b in bs]))
        assert self.As.ndim == 3 and self.bs.shape == self.As.shape[:2]
        self.rho, self.max_iter, self.threshold = float(rho), int(max_iter), float(threshold)
