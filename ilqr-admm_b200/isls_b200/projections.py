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


class SetConvexSOCRows:
    """State-side projection of ADMM_SLS (notebooks/Double integrator/LQR and SLS with state bounds.ipynb cell 16): each
    listed row of [d_x | Phi_x(:, :n/2)] (index into N * x_dim, negative indices allowed) is projected by its own
    `project_set_convex(x[row:row+1], A_, b_row, [project_soc_unit]*P, rho, max_iter, threshold)`; all other rows pass
    through.  The cone matrices A_ and the inner parameters are those of the control-side `SetConvexSOC`."""

    def __init__(self, rows, bs):
        self.rows = [int(r) for r in rows]
        self.bs = np.ascontiguousarray(np.stack([np.stack([np.asarray(b, dtype=np.float64) for b in row_bs])
                                                 for row_bs in bs]))
        assert self.bs.ndim == 3 and self.bs.shape[0] == len(self.rows)


class SetConvexSOCComponents:
    """State-side projection of the robust iSLS-ADMM (`iSLS.isls_admm(project_x=...)`, isls/isls.py:631-638): the device
    form of the closure

        def project_x(x, x_nom):              # x: [N * x_dim, dim + 1] = [d_x | Phi_x(:, :dim)]
            y = x.copy(); y[:, 0] += x_nom.flatten()
            for g, comp in enumerate(components):
                rows = np.arange(N) * x_dim + comp
                y[rows] = project_set_convex(y[rows], As, bs[g], [project_soc_unit] * P, rho, max_iter, threshold)
            y[:, 0] -= x_nom.flatten(); return y

    i.e. chance-constrained bounds on the listed state components (each with its own cone offsets bs[g], shared cone
    matrices As), all other rows passed through.  As: list of P [dim + 2, dim + 1]; bs: [G][P][dim + 2]."""

    def __init__(self, components, As, bs, rho=1.0, max_iter=200, threshold=1e-4):
        self.components = [int(c) for c in components]
        self.As = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in As]))
        self.bs = np.ascontiguousarray(np.asarray(bs, dtype=np.float64))
        assert self.As.ndim == 3 and self.bs.shape == (len(self.components),) + self.As.shape[:2]
        self.rho, self.max_iter, self.threshold = float(rho), int(max_iter), float(threshold)


# ---------------------------------------------------------------- batched row projections on the device (CUDA tensors)
def _rows(kind, x, p0=None, p1=None, l=0.0, u=0.0):
    import ctypes as C
    from . import _lib
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.ndim == 2):
        raise TypeError("device projections take a CUDA float64 tensor [rows, dim]")
    x = x.contiguous()
    out = torch.empty_like(x)
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float64) if not isinstance(a, torch.Tensor)
                                                         else a, dtype=torch.float64).to(x.device).expand(x.shape[1]).contiguous()
    p0, p1 = t(p0), t(p1)
    ptr = lambda a: None if a is None else C.c_void_p(a.data_ptr())
    with torch.cuda.device(x.device):
        rc = _lib.lib().isls_project_rows_f64(kind, x.shape[0], x.shape[1], ptr(x), ptr(p0), ptr(p1), float(l), float(u),
                                              ptr(out), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc, "isls_project_rows_f64")
    return out


def project_bound_batch(x, l, u):
    """isls/projections.py:7-11 on rows of a CUDA tensor (per-component bounds)."""
    return _rows(0, x, l, u)


def project_linear_batch(x, a, l, u):
    """isls/projections.py:30-43."""
    return _rows(1, x, a, None, l, u)


def project_quadratic_batch(x, l, u, center=None):
    """isls/projections.py:86-104; `center` c evaluates project_quadratic(x - c, l, u) + c (obstacle notebooks)."""
    return _rows(2, x, center, None, l, u)


def project_quadratic_b_batch(x, b, l, u):
    """isls/projections.py:106-115: l <= 0.5 x'x + b'x <= u."""
    b = np.asarray(b, dtype=np.float64)
    const = 0.5 * float(b @ b)
    return _rows(2, x, -b, None, l + const, u + const)


def project_soc_unit_batch(x):
    """isls/projections.py:140-162 on rows [z, t] (numpy batch semantics incl. SURVEY D9)."""
    return _rows(3, x)


def project_square_batch(x, l, u, center=None):
    """isls/projections.py:252-272 (`center`: project_square_c)."""
    return _rows(4, x, center, None, l, u)


def project_unit_ball_batch(x):
    """isls/projections.py:232-240 row-wise."""
    return _rows(5, x)


def project_affine_batch(x, a, b, l, u):
    """isls/projections.py:64-68: l <= a'x + b <= u, i.e. project_linear with the bounds shifted by b."""
    return _rows(1, x, a, None, l - b, u - b)


def _rows_ex(kind, x, A=None, b=None, l=None, u=None, rho=1.0, tol=1e-5, max_iter=100, blt=None, want_iters=False):
    import ctypes as C
    from . import _lib
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.ndim == 2):
        raise TypeError("device projections take a CUDA float64 tensor [rows, dim]")
    keep = {}

    def hp(a, shape):
        if a is None:
            return None
        keep[len(keep)] = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), shape))
        return keep[len(keep) - 1].ctypes.data
    p = _lib.ProjParams(kind=kind, rho=float(rho), tol=float(tol), max_iter=int(max_iter))
    it = torch.zeros(1, dtype=torch.int32, device=x.device) if want_iters else None
    if blt is not None:
        p.x_dim, p.u_dim, p.N = blt
        out = x.contiguous().clone()
        rows, dim = 1, 1
    else:
        A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        p.k = A.shape[0]
        p.A, p.b, p.l, p.u = hp(A, A.shape), hp(b, (p.k,)), hp(l, (p.k,)), hp(u, (p.k,))
        x = x.contiguous()
        out = torch.empty_like(x)
        rows, dim = x.shape
    with torch.cuda.device(x.device):
        rc = _lib.lib().isls_project_rows_ex_f64(C.byref(p), rows, dim, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                                 None if it is None else C.c_void_p(it.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc, "isls_project_rows_ex_f64")
    return (out, int(it[0])) if want_iters else out


def project_multilinear_batch(x, A, l, u):
    """isls/projections.py:46-62 on every row: x - A'(A A')^-1 (A x - clip(A x, l, u)) (a boundary projection, not the
    minimum-norm one - as the reference notes)."""
    return _rows_ex(6, x, A, None, l, u)


def project_soc_batch(z0, A, b, rho=1.0, max_iter=100, tol=1e-5, want_iters=False):
    """isls/projections.py:163-232: rows z0 onto {z : A z + b in SOC} by the reference's inner ADMM (stop rule: maximum of
    the residual norms over all rows, so at most 1,024 rows per call)."""
    return _rows_ex(7, z0, A, b, None, None, rho=rho, tol=tol, max_iter=max_iter, want_iters=want_iters)


def project_block_lower_triangular(z, x_dim, u_dim, N):
    """isls/projections.py:277-286 on a CUDA matrix [N u_dim, N x_dim]: zeroes z[i u_dim, i x_dim:(i+1) x_dim] (exactly
    the rows the reference touches).  Returns a new tensor."""
    if z.shape != (N * u_dim, N * x_dim):
        raise ValueError("z must be [N u_dim, N x_dim]")
    return _rows_ex(8, z, blt=(x_dim, u_dim, N))


_SET_KINDS = {"bound": 0, "quadratic": 2, "soc_unit": 3, "square": 4, "unit_ball": 5}


def project_set_convex_batch(x0, As, bs, projections, rho=1.0, max_iter=200, threshold=1e-4, want_iters=False):
    """isls/projections.py:289-374 on the device: rows x0 [rows <= 1024, dim] onto the intersection of the sets
    {x : A_i x + b_i in C_i}.  `projections[i]` describes C_i by one of the primitive batch projections instead of a Python
    callable: ("bound", lo, hi), ("quadratic", l, u[, center]), ("soc_unit",), ("square", l, u[, center]), ("unit_ball",)."""
    import ctypes as C
    from . import _lib
    if not (isinstance(x0, torch.Tensor) and x0.is_cuda and x0.dtype == torch.float64 and x0.ndim == 2):
        raise TypeError("device projections take a CUDA float64 tensor [rows, dim]")
    if not (len(As) == len(bs) == len(projections)) or not 1 <= len(As) <= 4:
        raise ValueError("1..4 sets, one (A, b, projection) each")
    x0 = x0.contiguous()
    dim = x0.shape[1]
    keep = []

    def hp(a, n):
        keep.append(np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,))))
        return keep[-1].ctypes.data
    p = _lib.ProjSetParams(n_sets=len(As), rho=float(rho), threshold=float(threshold), max_iter=int(max_iter))
    for i, (A, b, pr) in enumerate(zip(As, bs, projections)):
        A = np.ascontiguousarray(np.atleast_2d(np.asarray(A, dtype=np.float64)))
        if A.shape[1] != dim:
            raise ValueError("A_%d must have %d columns" % (i, dim))
        keep.append(A)
        e = p.sets[i]
        e.kind, e.rows, e.A, e.b = _SET_KINDS[pr[0]], A.shape[0], A.ctypes.data, hp(b, A.shape[0])
        if pr[0] == "bound":
            e.p0, e.p1 = hp(pr[1], A.shape[0]), hp(pr[2], A.shape[0])
        elif pr[0] in ("quadratic", "square"):
            e.l, e.u = float(pr[1]), float(pr[2])
            if len(pr) > 3 and pr[3] is not None:
                e.p0 = hp(pr[3], A.shape[0])
    out = torch.empty_like(x0)
    it = torch.zeros(1, dtype=torch.int32, device=x0.device)
    with torch.cuda.device(x0.device):
        rc = _lib.lib().isls_project_set_convex_f64(C.byref(p), x0.shape[0], dim, C.c_void_p(x0.data_ptr()),
                                                    C.c_void_p(out.data_ptr()), C.c_void_p(it.data_ptr()),
                                                    C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream))
    _lib.check(rc, "isls_project_set_convex_f64")
    return (out, int(it[0])) if want_iters else out


class ObstacleSets:
    """Device descriptor of the notebooks' obstacle-avoidance state projection (Car/Iterative LQR with state
    constraints.ipynb cell 18): `project_set_convex(x, [I]*K, [0]*K, projections, rho, max_iter, threshold)` where
    projection k maps the position p = x[:2] of every time step to  c_k + W_k^-1 Pi_sq(W_k (p - c_k))  with
    Pi_sq = project_square_batch(., lower_k, upper) (isls/projections.py:246-255, 289-374).  Pass it as `project_x`."""

    def __init__(self, centers, W=None, lower=None, upper=1e5, rho=1e1, max_iter=15, threshold=1e-3, kind="square",
                 dykstra_max_iter=0, dykstra_tol=1e-5):
        """kind="square": rotated infinity-norm shells (W required; all state components go through the consensus
        ADMM, As = I_n).  kind="quadratic": spherical shells lower <= 0.5 ||p - c||^2 <= upper on the position only
        (`project_quadratic`, isls/projections.py:91-105), followed by `project_set_convex_dykstra` when
        dykstra_max_iter > 0 - the project_state of Double integrator/LQR and SLS with spherical obstacle
        avoidance.ipynb cell 12 (LQT path: SLS.ADMM_LQT_DP / ADMM_LQT_Batch)."""
        self.centers = np.asarray(centers, dtype=np.float64).reshape(-1, 2)
        K = self.centers.shape[0]
        self.kind = kind
        if kind not in ("square", "quadratic"):
            raise ValueError("kind must be 'square' or 'quadratic'")
        self.W = np.tile(np.eye(2), (K, 1, 1)) if W is None else np.asarray(W, dtype=np.float64).reshape(K, 2, 2)
        self.W_inv = np.linalg.inv(self.W)
        self.lower = np.asarray(lower, dtype=np.float64).reshape(K)
        self.upper, self.rho, self.max_iter, self.threshold = float(upper), float(rho), int(max_iter), float(threshold)
        self.dykstra_max_iter, self.dykstra_tol = int(dykstra_max_iter), float(dykstra_tol)

    def as_dict(self):
        return dict(kind=self.kind, centers=self.centers, W=self.W, W_inv=self.W_inv, lower=self.lower,
                    upper=self.upper, rho=self.rho, max_iter=self.max_iter, threshold=self.threshold,
                    dykstra_max_iter=self.dykstra_max_iter, dykstra_tol=self.dykstra_tol)

    def key(self):
        return (self.kind, self.centers.tobytes(), self.W.tobytes(), self.lower.tobytes(), self.upper, self.rho,
                self.max_iter, self.threshold, self.dykstra_max_iter, self.dykstra_tol)
