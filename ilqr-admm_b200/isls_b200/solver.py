"""Thin host layer over the C-ABI: plans, workspaces and batched solves on torch CUDA tensors.

torch is used only as the buffer carrier (allocation, streams, host<->device copies); all arithmetic of the hot
path runs in libisls_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

MODEL_DIMS = {"car": (4, 2), "arm3": (9, 3), "tassa_car": (4, 2)}
COST_KINDS = {"quadratic": 0, "pseudo_huber": 1}

ST_CONVERGED_COST, ST_LINESEARCH_FAIL, ST_MAX_ITER, ST_OSCILLATING, ST_NON_PD, ST_NAN_COST = 1, 2, 4, 8, 16, 32
ADMM_CONVERGED, ADMM_STALLED, ADMM_MAXIT = 1, 2, 3


def alphas(L):
    """Line-search step sizes of the reference (isls/isls_base.py:10-11)."""
    return (10.0 ** np.linspace(0.0, -5.0, 50))[:L].copy()


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Plan:
    """Problem description shared by a batch (isls_problem_desc): model, horizon, quadratic via-point cost,
    diagonal ADMM penalties and box bounds."""

    def __init__(self, model, N, n, m, dt, Qdiag, seq, u_std, L, rho_x=None, lo_x=None, hi_x=None, rho_u=None,
                 lo_u=None, hi_u=None, cost="quadratic", Rdiag=None, Hp=None, Qdiag_b=None, Hp_b=None, obstacles=None, isls_dim=0,
                 device="cuda:0", lti_AB=None):
        """cost="pseudo_huber": Qdiag / Hp are the weights and smoothness scales [n_via, n] of the first term,
        Qdiag_b / Hp_b of the optional second one (Tutorial cell 14); Rdiag [m] replaces R = u_std I.
        device: the GPU that holds the plan's constant block - a BatchSolver must use the same one."""
        L_ = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.IslsError("isls_b200 needs a CUDA device (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        mid = L_.isls_model_id(model.encode())
        _lib.check(0 if mid >= 0 else mid, "isls_model_id(%r)" % model)
        _lib.check(L_.isls_model_supported(mid, n, m), "isls_model_supported(%s, n=%d, m=%d)" % (model, n, m))
        self.model, self.N, self.n, self.m, self.dt, self.L = model, int(N), int(n), int(m), float(dt), int(L)
        Qdiag = _f64(Qdiag)
        self.n_via = Qdiag.shape[0]
        keep = dict(Qdiag=Qdiag.reshape(self.n_via, n), seq=np.ascontiguousarray(seq, dtype=np.int32),
                    alphas=alphas(L))
        self.proj_x = rho_x is not None
        self.proj_u = rho_u is not None
        if obstacles is not None and not self.proj_x:
            raise _lib.IslsError("obstacle sets are a state projection: rho_x is required")
        if self.proj_x and obstacles is not None:
            keep.update(rho_x=_f64(rho_x, (N, n)))
        elif self.proj_x:
            keep.update(rho_x=_f64(rho_x, (N, n)), lo_x=_f64(-np.inf if lo_x is None else lo_x, (N, n)),
                        hi_x=_f64(np.inf if hi_x is None else hi_x, (N, n)))
        if self.proj_u:
            keep.update(rho_u=_f64(rho_u, (N, m)), lo_u=_f64(-np.inf if lo_u is None else lo_u, (N, m)),
                        hi_u=_f64(np.inf if hi_u is None else hi_u, (N, m)))
        assert keep["seq"].shape == (N,)
        if cost not in COST_KINDS:
            raise _lib.IslsError("unknown cost %r (device costs: %s)" % (cost, tuple(COST_KINDS)))
        for nm, arr, shp in (("Rdiag", Rdiag, (m,)), ("Hp", Hp, (self.n_via, n)), ("Qdiag_b", Qdiag_b, (self.n_via, n)),
                             ("Hp_b", Hp_b, (self.n_via, n))):
            if arr is not None:
                keep[nm] = _f64(arr, shp)
        d = _lib.ProblemDesc(model_id=mid, n=n, m=m, N=N, n_via=self.n_via, L=L, dt=dt, u_std=float(u_std),
                             cost_kind=COST_KINDS[cost], isls_dim=int(isls_dim))
        self.isls_dim = int(isls_dim)
        if obstacles is not None:
            ob = obstacles
            K = len(ob["centers"])
            keep.update(obst_centers=_f64(ob["centers"], (K, 2)), obst_W=_f64(ob["W"], (K, 2, 2)),
                        obst_W_inv=_f64(ob["W_inv"], (K, 2, 2)), obst_lower=_f64(ob["lower"], (K,)))
            d.n_obst, d.obst_max_iter = K, int(ob["max_iter"])
            d.obst_kind = {"square": 0, "quadratic": 1}[ob.get("kind", "square")]
            d.obst_dykstra_max_iter, d.obst_dykstra_tol = int(ob.get("dykstra_max_iter", 0)), float(ob.get("dykstra_tol", 0.0))
            d.obst_upper, d.obst_rho, d.obst_threshold = float(ob["upper"]), float(ob["rho"]), float(ob["threshold"])
            for k in ("obst_centers", "obst_W", "obst_W_inv", "obst_lower"):
                setattr(d, k, _ptr(keep[k]))
        if lti_AB is not None:            # model "lti": x+ = A x + B u with the given constant matrices
            keep.update(lti_A=_f64(lti_AB[0], (n, n)), lti_B=_f64(lti_AB[1], (n, m)))
        for k in ("Qdiag", "seq", "alphas", "rho_x", "lo_x", "hi_x", "rho_u", "lo_u", "hi_u", "Rdiag", "Hp", "Qdiag_b",
                  "Hp_b", "lti_A", "lti_B"):
            setattr(d, k, _ptr(keep.get(k)))
        self._keep = keep
        h = C.c_void_p()
        with torch.cuda.device(self.device):          # the constant block is allocated on the current device
            _lib.check(L_.isls_plan_create(C.byref(d), C.byref(h)), "isls_plan_create")
        self.handle = h

    def workspace_bytes(self, B):
        sz = C.c_size_t()
        _lib.check(_lib.lib().isls_workspace_bytes(self.handle, int(B), C.byref(sz)), "isls_workspace_bytes")
        return sz.value

    def close(self):
        if getattr(self, "handle", None):
            with torch.cuda.device(self.device):
                _lib.lib().isls_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Result(dict):
    __getattr__ = dict.__getitem__


class BatchSolver:
    """Owns the workspace and the result buffers for a fixed (plan, B); solves are enqueued on the current
    torch stream.  Inputs may be host (numpy / pinned torch) or device tensors."""

    def __init__(self, plan, B, device="cuda:0", max_outer=20, max_admm=20, logs=True, want_gains=False,
                 want_masks=False, want_Qs=False):
        if not torch.cuda.is_available():
            raise _lib.IslsError("isls_b200 needs a CUDA device (there is no CPU path)")
        self.plan, self.B, self.device = plan, int(B), torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self.device != plan.device:
            raise _lib.IslsError("plan constants live on %s but the solver was asked for %s" % (plan.device, self.device))
        self.max_outer, self.max_admm = int(max_outer), int(max_admm)
        p, dev = plan, self.device
        with torch.cuda.device(dev):
            self.ws = torch.empty(plan.workspace_bytes(B) + 256, dtype=torch.uint8, device=dev)
        off = (-self.ws.data_ptr()) % 256
        self._ws_ptr, self._ws_bytes = self.ws.data_ptr() + off, self.ws.numel() - off
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        B_, N, n, m = self.B, p.N, p.n, p.m
        o = Result(x=torch.empty(B_, N, n, **f64), u=torch.empty(B_, N, m, **f64), cost=torch.empty(B_, **f64),
                   cost_log=torch.empty(B_, self.max_outer + 1, **f64), n_log=torch.empty(B_, **i32),
                   status=torch.empty(B_, **i32), outer_iters=torch.empty(B_, **i32))
        if logs:
            o.update(admm_iters=torch.empty(B_, self.max_outer, **i32),
                     admm_exit=torch.empty(B_, self.max_outer, **i32),
                     res_log=torch.empty(B_, self.max_outer, max(self.max_admm, 1), 2, **f64),
                     alpha_idx=torch.empty(B_, self.max_outer, max(self.max_admm, 1), **i32),
                     inner_iters=torch.empty(B_, self.max_outer, max(self.max_admm, 1), **i32),
                     z_x=torch.empty(B_, N, n, **f64), z_u=torch.empty(B_, N, m, **f64),
                     lam_x=torch.empty(B_, N, n, **f64), lam_u=torch.empty(B_, N, m, **f64))
        if want_gains:
            o.update(K=torch.empty(B_, N, m, n, **f64), k=torch.empty(B_, N, m, **f64))
        if want_Qs:               # Riccati logs of SLS.solve_dp(return_Qs=True) (isls/sls.py:117-120)
            o.update(Quu=torch.empty(B_, N, m, m, **f64), Quu_inv=torch.empty(B_, N, m, m, **f64),
                     Qux=torch.empty(B_, N, m, n, **f64))
        if want_masks:
            o.update(mask_x=torch.zeros(B_, N, n, dtype=torch.int8, device=dev),
                     mask_u=torch.zeros(B_, N, m, dtype=torch.int8, device=dev))
        self._sets = [(o, self._make_cout(o))]
        self.out, self._cout = self._sets[0]
        self.h2d_bytes = 0
        # device staging buffers for the inputs
        self.x0 = torch.empty(B_, n, **f64)
        self.u_init = torch.empty(B_, N, m, **f64)
        self.zs = torch.empty(B_, p.n_via, n, **f64)

    @staticmethod
    def _make_cout(o):
        c = _lib.SolveOut()
        for f in _lib.OUT_FIELDS:
            setattr(c, f, _dptr(o.get(f)))
        return c

    def add_output_set(self):
        """Second (third, ...) set of result buffers: a caller that copies results to the host on a side stream while the
        next solve runs alternates between sets with use_output_set() (bench.py's end-to-end arm)."""
        o = Result({k: torch.empty_like(v) for k, v in self._sets[0][0].items()})
        self._sets.append((o, self._make_cout(o)))
        return len(self._sets) - 1

    def use_output_set(self, i):
        self.out, self._cout = self._sets[i]

    # ---- input staging (host -> device copies happen here when host arrays are given)
    def _stage(self, dst, src, name):
        if isinstance(src, torch.Tensor):
            t = src
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(src, dtype=np.float64)))
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        if t.device != dst.device:
            # host -> device: move the bytes that exist (un-expanded storage), broadcast on the device
            base = t
            if any(st == 0 for st in t.stride()) and t.numel() > 0:
                idx = tuple(slice(0, 1) if st == 0 else slice(None) for st in t.stride())
                base = t[idx]
            self.h2d_bytes += base.numel() * 8
            t = base.to(dst.device, non_blocking=True)
        dst.copy_(t.expand(dst.shape), non_blocking=True)

    def set_inputs(self, x0, u_init, zs):
        self._stage(self.x0, x0, "x0")
        self._stage(self.u_init, u_init, "u_init")
        self._stage(self.zs, zs, "zs")

    def _opts(self, tol, outer_tol, relax, fixed_budget, last_stage_dp, max_outer=None, max_admm=None, stall_tol=0.0,
              osc_tol=0.0):
        return _lib.SolveOpts(max_outer=self.max_outer if max_outer is None else max_outer,
                              max_admm=self.max_admm if max_admm is None else max_admm, tol=tol,
                              outer_tol=outer_tol, relax=relax, fixed_budget=int(fixed_budget),
                              last_stage_dp=int(last_stage_dp), stall_tol=stall_tol, osc_tol=osc_tol)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def ilqr_admm(self, tol=1e-3, outer_tol=1e-3, relax=1.0, fixed_budget=False, last_stage_dp=False):
        o = self._opts(tol, outer_tol, relax, fixed_budget, last_stage_dp)
        with torch.cuda.device(self.device):
            rc = _lib.lib().isls_ilqr_admm_solve_f64(self.plan.handle, C.byref(o), self.B, _dptr(self.x0),
                                                     _dptr(self.u_init), _dptr(self.zs), C.c_void_p(self._ws_ptr),
                                                     self._ws_bytes, C.byref(self._cout), self._stream())
        _lib.check(rc, "isls_ilqr_admm_solve_f64")
        return self.out

    def probe_overlap(self, ls_ctas=2, ff_depth=4):
        """Measurement helper (isls_probe_overlap_f64): ms of the line search / ff-pass of the two half batches alone and
        together.  -> dict"""
        o = self._opts(1e-3, 1e-3, 1.0, True, False)
        ms = (C.c_double * 6)()
        with torch.cuda.device(self.device):
            rc = _lib.lib().isls_probe_overlap_f64(self.plan.handle, C.byref(o), self.B, _dptr(self.x0), _dptr(self.u_init),
                                                   _dptr(self.zs), C.c_void_p(self._ws_ptr), self._ws_bytes,
                                                   C.byref(self._cout), int(ls_ctas), int(ff_depth), ms, self._stream())
        _lib.check(rc, "isls_probe_overlap_f64")
        return dict(zip(("ls_plain", "ls_pers", "ff_tma", "ff_plain", "ls_pers||ff_tma", "ls_plain||ff_tma"), list(ms)))

    def isls_admm(self, soc, tol=1e-3, relax=1.0, fixed_budget=False, soc_x=None):
        """Robust iSLS-ADMM (isls_isls_admm_solve_f64; isls/isls.py:503-712).  soc: projections.SetConvexSOC (control
        side) or None; soc_x: projections.SetConvexSOCComponents (state side; the plan then needs rho_x) or None.
        Adds d_u [B,N,m] and phi_u [B,N,m,dim] to the results."""
        p, dev = self.plan, self.device
        if p.isls_dim < 1:
            raise _lib.IslsError("the plan was not created with isls_dim > 0")
        C_ = p.isls_dim + 1
        o = self._opts(tol, 1e-4, relax, fixed_budget, False, stall_tol=1e-3, osc_tol=1e-3)   # isls.py:664, 700, 704
        if soc_x is not None:
            if soc_x.As.shape[1:] != (C_ + 1, C_):
                raise ValueError("cone matrices must be [%d, %d] (dim + 2 rows, dim + 1 columns)" % (C_ + 1, C_))
            if soc is not None and not (np.array_equal(soc.As, soc_x.As) and (soc.rho, soc.max_iter, soc.threshold) ==
                                        (soc_x.rho, soc_x.max_iter, soc_x.threshold)):
                raise ValueError("the state and control sides share the cone matrices and the inner ADMM parameters")
            comps = np.ascontiguousarray(np.asarray(soc_x.components, dtype=np.int32))
            so = _lib.SlsAdmmOpts(max_iter=0, rho_u=0.0, alpha=relax, tol=tol, fixed_budget=int(fixed_budget),
                                  n_cones=soc_x.As.shape[0], cone_rows=C_ + 1, As=soc_x.As.ctypes.data,
                                  bs=soc.bs.ctypes.data if soc is not None else None, inner_rho=soc_x.rho,
                                  inner_max_iter=soc_x.max_iter, inner_threshold=soc_x.threshold,
                                  n_x_rows=len(comps), x_row_idx=comps.ctypes.data, x_bs=soc_x.bs.ctypes.data)
        elif soc is None:                    # no projection: unconstrained iSLS step (z = x)
            so = _lib.SlsAdmmOpts(max_iter=0, rho_u=0.0, alpha=relax, tol=tol, fixed_budget=int(fixed_budget),
                                  n_cones=0, cone_rows=C_ + 1, inner_rho=1.0, inner_max_iter=1, inner_threshold=1.0)
        else:
            if soc.As.shape[1:] != (C_ + 1, C_):
                raise ValueError("cone matrices must be [%d, %d] (dim + 2 rows, dim + 1 columns)" % (C_ + 1, C_))
            so = _lib.SlsAdmmOpts(max_iter=0, rho_u=0.0, alpha=relax, tol=tol, fixed_budget=int(fixed_budget),
                                  n_cones=soc.As.shape[0], cone_rows=C_ + 1, As=soc.As.ctypes.data,
                                  bs=soc.bs.ctypes.data, inner_rho=soc.rho, inner_max_iter=soc.max_iter,
                                  inner_threshold=soc.threshold)
        if "d_u" not in self.out:
            f64 = dict(dtype=torch.float64, device=dev)
            self.out.update(d_u=torch.empty(self.B, p.N, p.m, **f64), phi_u=torch.empty(self.B, p.N, p.m, p.isls_dim, **f64))
        with torch.cuda.device(dev):
            rc = _lib.lib().isls_isls_admm_solve_f64(p.handle, C.byref(o), C.byref(so), self.B, _dptr(self.x0),
                                                     _dptr(self.u_init), _dptr(self.zs), C.c_void_p(self._ws_ptr),
                                                     self._ws_bytes, C.byref(self._cout), _dptr(self.out.d_u),
                                                     _dptr(self.out.phi_u), self._stream())
        _lib.check(rc, "isls_isls_admm_solve_f64")
        return self.out

    def ilqr(self, tol_fun=1e-5, fixed_budget=False):
        o = self._opts(tol_fun, 0.0, 1.0, fixed_budget, True)
        with torch.cuda.device(self.device):
            rc = _lib.lib().isls_ilqr_solve_f64(self.plan.handle, C.byref(o), self.B, _dptr(self.x0),
                                                _dptr(self.u_init), _dptr(self.zs), C.c_void_p(self._ws_ptr),
                                                self._ws_bytes, C.byref(self._cout), self._stream())
        _lib.check(rc, "isls_ilqr_solve_f64")
        return self.out

    def lqt_admm_dp(self, tol=1e-3, relax=1.0, fixed_budget=False, last_stage_dp=True, z_x_init=None, z_u_init=None):
        """z_x_init [B,N,n] / z_u_init [B,N,m]: device tensors (ADMM warm start) or None."""
        o = self._opts(tol, 0.0, relax, fixed_budget, last_stage_dp, max_outer=1)
        self._zinit = (None if z_x_init is None else z_x_init.contiguous(),
                       None if z_u_init is None else z_u_init.contiguous())       # keep alive until the launch
        o.z_x_init_dev, o.z_u_init_dev = (None if t is None else t.data_ptr() for t in self._zinit)
        with torch.cuda.device(self.device):
            rc = _lib.lib().isls_lqt_admm_dp_f64(self.plan.handle, C.byref(o), self.B, _dptr(self.x0),
                                                 _dptr(self.zs), C.c_void_p(self._ws_ptr), self._ws_bytes,
                                                 C.byref(self._cout), self._stream())
        _lib.check(rc, "isls_lqt_admm_dp_f64")
        return self.out

    def linesearch(self, x_nom, u_nom, du, zs, reg_x=None, reg_u=None):
        """Stage-level: open-loop line search + argmin (isls_rollout_linesearch_f64)."""
        p, dev = self.plan, self.device
        f64 = dict(dtype=torch.float64, device=dev)

        def dv(a, shape):
            if a is None:
                return None
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
            return t.to(dev).expand(shape).contiguous()
        B_, N, n, m = self.B, p.N, p.n, p.m
        x_nom, u_nom, du = dv(x_nom, (B_, N, n)), dv(u_nom, (B_, N, m)), dv(du, (B_, N, m))
        zs, reg_x, reg_u = dv(zs, (B_, p.n_via, n)), dv(reg_x, (B_, N, n)), dv(reg_u, (B_, N, m))
        costs = torch.empty(B_, p.L, **f64)
        best = torch.empty(B_, dtype=torch.int32, device=dev)
        xb, ub = torch.empty(B_, N, n, **f64), torch.empty(B_, N, m, **f64)
        with torch.cuda.device(dev):
            rc = _lib.lib().isls_rollout_linesearch_f64(p.handle, B_, _dptr(x_nom), _dptr(u_nom), _dptr(du),
                                                        _dptr(zs), _dptr(reg_x), _dptr(reg_u), _dptr(costs),
                                                        _dptr(best), _dptr(xb), _dptr(ub), C.c_void_p(self._ws_ptr),
                                                        self._ws_bytes, self._stream())
        _lib.check(rc, "isls_rollout_linesearch_f64")
        return costs, best, xb, ub


def riccati(A, Bm, c, Cm):
    """Stage-level generic-operator Riccati pass (isls_riccati_f64; iSLS.backward_pass_DP, isls/isls.py:229-308).
    A[B,N,n,n], Bm[B,N,n,m], c[B,N,n+m], Cm[B,N,n+m,n+m] (CUDA float64) -> K[B,N,m,n], k[B,N,m], non_pd[B]."""
    B_, N, n, m = Bm.shape
    dev = A.device
    A, Bm, c, Cm = (t.contiguous() for t in (A, Bm, c, Cm))
    K = torch.empty(B_, N, m, n, dtype=torch.float64, device=dev)
    k = torch.empty(B_, N, m, dtype=torch.float64, device=dev)
    bad = torch.empty(B_, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().isls_riccati_f64(n, m, N, B_, _dptr(A), _dptr(Bm), _dptr(c), _dptr(Cm), _dptr(K), _dptr(k),
                                         _dptr(bad), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "isls_riccati_f64")
    return K, k, bad


def admm_project_dual(x, z, lam, lo, hi, relax=1.0, want_mask=False):
    """Stage-level ADMM z-projection + scaled dual update on flat [B,len] CUDA tensors, in place on z, lam
    (isls_admm_project_dual_f64; isls/admm.py:43-69 with project_bound).  Returns (prim_sq[B], dual_sq[B], mask)."""
    B_, ln = x.shape
    dev = x.device
    prim = torch.zeros(B_, dtype=torch.float64, device=dev)
    dual = torch.zeros(B_, dtype=torch.float64, device=dev)
    mask = torch.empty(B_, ln, dtype=torch.int8, device=dev) if want_mask else None
    lo = lo.to(dev).expand(ln).contiguous()
    hi = hi.to(dev).expand(ln).contiguous()
    assert x.is_contiguous() and z.is_contiguous() and lam.is_contiguous()
    with torch.cuda.device(dev):
        rc = _lib.lib().isls_admm_project_dual_f64(B_, ln, float(relax), _dptr(x), _dptr(z), _dptr(lam), _dptr(lo),
                                                   _dptr(hi), _dptr(prim), _dptr(dual), _dptr(mask),
                                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "isls_admm_project_dual_f64")
    return prim, dual, mask


def measure_fp64_tflops(device="cuda:0"):
    v = C.c_double()
    with torch.cuda.device(device):
        rc = _lib.lib().isls_measure_fp64_tflops(C.byref(v),
                                                 C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _lib.check(rc, "isls_measure_fp64_tflops")
    return v.value


def profile_enable(on):
    _lib.check(_lib.lib().isls_profile_enable(int(bool(on))), "isls_profile_enable")


def profile_collect():
    """-> {kernel class: (total ms, launches)} since the last collect (per-kernel CUDA-event timing)."""
    n = len(_lib.KERNEL_CLASSES)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    _lib.check(_lib.lib().isls_profile_collect(ms, cnt), "isls_profile_collect")
    return {k: (ms[i], cnt[i]) for i, k in enumerate(_lib.KERNEL_CLASSES) if cnt[i]}


def mc_rollout(model, n, m, N, dt, mode, x0, K, k, x_nom=None, u_nom=None, noise_scale=0.0, seed=0, device="cuda:0"):
    """Monte-Carlo closed-loop rollouts of one controller over the rows of x0 (isls_mc_rollout_f64).
    mode: "batch" (open loop, k = us[N,m]), "dp" (K[N,m,n], k[N,m]), "sls" (K[N m, N n], k[N m])."""
    L_ = _lib.lib()
    mid = L_.isls_model_id(model.encode())
    _lib.check(0 if mid >= 0 else mid, "isls_model_id(%r)" % model)
    dev = torch.device(device)

    def dv(a):
        if a is None:
            return None
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
        return t.to(dev, dtype=torch.float64).contiguous()
    x0, K, k, x_nom, u_nom = dv(x0), dv(K), dv(k), dv(x_nom), dv(u_nom)
    x0 = x0.reshape(-1, n)
    B_ = x0.shape[0]
    xs = torch.empty(B_, N, n, dtype=torch.float64, device=dev)
    us = torch.empty(B_, N, m, dtype=torch.float64, device=dev)
    md = {"batch": 0, "dp": 1, "sls": 2}[mode]
    with torch.cuda.device(dev):
        rc = L_.isls_mc_rollout_f64(mid, n, m, N, float(dt), md, B_, _dptr(x0), _dptr(K), _dptr(k), _dptr(x_nom),
                                    _dptr(u_nom), float(noise_scale), int(seed), _dptr(xs), _dptr(us),
                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "isls_mc_rollout_f64")
    return xs, us


def _dev_f64(a, dev):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
    return t.to(dev, dtype=torch.float64).contiguous()


def linearize(model, n, m, dt, x, u, device="cuda:0"):
    """A, B = linearize(...): the notebooks' `get_AB(x_nom, u_nom)` on the device (isls_linearize_f64) for x [..., n],
    u [..., m] -> A [..., n, n], B [..., n, m]."""
    L_ = _lib.lib()
    mid = L_.isls_model_id(model.encode())
    _lib.check(0 if mid >= 0 else mid, "isls_model_id(%r)" % model)
    dev = torch.device(device)
    x, u = _dev_f64(x, dev), _dev_f64(u, dev)
    lead = tuple(x.shape[:-1])
    if x.shape[-1] != n or u.shape[-1] != m or tuple(u.shape[:-1]) != lead:
        raise ValueError("x must be [..., %d] and u [..., %d] with the same leading shape" % (n, m))
    rows = int(np.prod(lead)) if lead else 1
    A = torch.empty(lead + (n, n), dtype=torch.float64, device=dev)
    Bm = torch.empty(lead + (n, m), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = L_.isls_linearize_f64(mid, n, m, float(dt), rows, _dptr(x), _dptr(u), _dptr(A), _dptr(Bm),
                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "isls_linearize_f64")
    return A, Bm


def controller_tv(A, Bm, PHI_U, du, device="cuda:0"):
    """K, k = PHI_U PHI_X^-1, (I - K D) du with PHI_X = C + D PHI_U and the time-varying operators C, D of A_t, B_t
    (isls_controller_tv_f64; isls/sls.py:235-242 on isls/isls_base.py:138-158).  A [N, n, n] / B [N, n, m] shared by the
    batch or [B, N, n, n] / [B, N, n, m]; PHI_U [B, N m, N n] (or one [N m, N n]), du [B, N m]."""
    L_ = _lib.lib()
    dev = torch.device(device)
    A, Bm, PHI_U, du = _dev_f64(A, dev), _dev_f64(Bm, dev), _dev_f64(PHI_U, dev), _dev_f64(du, dev)
    single = PHI_U.ndim == 2
    if single:
        PHI_U, du = PHI_U[None], du[None]
    shared = A.ndim == 3
    N, n, m = int(A.shape[-3]), int(A.shape[-1]), int(Bm.shape[-1])
    B_ = int(PHI_U.shape[0])
    if PHI_U.shape != (B_, N * m, N * n) or du.shape != (B_, N * m) or Bm.shape[-3:] != (N, n, m) or (
            not shared and (A.shape[0] != B_ or Bm.shape[0] != B_)):
        raise ValueError("shapes: A [(B,) N, n, n], B [(B,) N, n, m], PHI_U [B, N m, N n], du [B, N m]")
    # causal PHI_U only (u_t reacts to w_s, s <= t): the back-substitution relies on PHI_X being unit block lower
    # triangular; the reference's dense inverse would accept anything
    blk = torch.arange(N, device=dev)
    upper = (blk.repeat_interleave(m)[:, None] < blk.repeat_interleave(n)[None, :])
    if bool((PHI_U[:, upper] != 0).any()):
        raise ValueError("PHI_U must be block lower triangular (causal)")
    nbytes = L_.isls_controller_tv_workspace_bytes(n, m, N, B_)
    ws = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    K = torch.empty(B_, N * m, N * n, dtype=torch.float64, device=dev)
    k = torch.empty(B_, N * m, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = L_.isls_controller_tv_f64(n, m, N, B_, _dptr(A), _dptr(Bm), int(shared), _dptr(PHI_U), _dptr(du), _dptr(ws),
                                       nbytes, _dptr(K), _dptr(k), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "isls_controller_tv_f64")
    return (K[0], k[0]) if single else (K, k)
