"""ctypes binding of libisls_b200.so (the C-ABI declared in include/isls_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libisls_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class _Guarded(C.Structure):
    """POD struct of the C-ABI: the first field is the ABI guard `struct_size`, filled in automatically."""

    def __init__(self, *a, **kw):
        super().__init__(**kw)
        if a:
            raise TypeError("keyword arguments only")
        self.struct_size = C.sizeof(self)


class ProblemDesc(_Guarded):
    _fields_ = [("struct_size", C.c_uint32), ("model_id", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("N", C.c_int32),
                ("n_via", C.c_int32), ("L", C.c_int32), ("dt", C.c_double), ("u_std", C.c_double),
                ("Qdiag", C.c_void_p), ("seq", C.c_void_p), ("alphas", C.c_void_p),
                ("rho_x", C.c_void_p), ("lo_x", C.c_void_p), ("hi_x", C.c_void_p),
                ("rho_u", C.c_void_p), ("lo_u", C.c_void_p), ("hi_u", C.c_void_p),
                ("cost_kind", C.c_int32), ("Rdiag", C.c_void_p), ("Hp", C.c_void_p), ("Qdiag_b", C.c_void_p),
                ("Hp_b", C.c_void_p), ("n_obst", C.c_int32), ("obst_max_iter", C.c_int32),
                ("obst_centers", C.c_void_p), ("obst_W", C.c_void_p), ("obst_W_inv", C.c_void_p),
                ("obst_lower", C.c_void_p), ("obst_upper", C.c_double), ("obst_rho", C.c_double),
                ("obst_threshold", C.c_double), ("obst_kind", C.c_int32), ("obst_dykstra_max_iter", C.c_int32),
                ("obst_dykstra_tol", C.c_double), ("isls_dim", C.c_int32), ("lti_A", C.c_void_p), ("lti_B", C.c_void_p)]


class SolveOpts(_Guarded):
    _fields_ = [("struct_size", C.c_uint32), ("max_outer", C.c_int32), ("max_admm", C.c_int32), ("tol", C.c_double), ("outer_tol", C.c_double),
                ("relax", C.c_double), ("fixed_budget", C.c_int32), ("last_stage_dp", C.c_int32),
                ("stall_tol", C.c_double), ("osc_tol", C.c_double), ("z_x_init_dev", C.c_void_p),
                ("z_u_init_dev", C.c_void_p)]


OUT_FIELDS = ["x", "u", "cost", "cost_log", "n_log", "status", "outer_iters", "admm_iters", "admm_exit", "res_log",
              "alpha_idx", "z_x", "z_u", "lam_x", "lam_u", "K", "k", "mask_x", "mask_u", "inner_iters", "Quu", "Quu_inv", "Qux"]


class SlsAdmmOpts(_Guarded):
    _fields_ = [("struct_size", C.c_uint32), ("max_iter", C.c_int32), ("rho_u", C.c_double), ("alpha", C.c_double), ("tol", C.c_double),
                ("fixed_budget", C.c_int32), ("n_cones", C.c_int32), ("cone_rows", C.c_int32), ("As", C.c_void_p),
                ("bs", C.c_void_p), ("inner_rho", C.c_double), ("inner_max_iter", C.c_int32),
                ("inner_threshold", C.c_double), ("n_x_rows", C.c_int32), ("x_row_idx", C.c_void_p),
                ("x_bs", C.c_void_p), ("rho_x_rows", C.c_void_p)]


class ProjParams(_Guarded):
    _fields_ = [("struct_size", C.c_uint32), ("kind", C.c_int32), ("k", C.c_int32), ("A", C.c_void_p), ("b", C.c_void_p),
                ("l", C.c_void_p), ("u", C.c_void_p), ("rho", C.c_double), ("tol", C.c_double), ("max_iter", C.c_int32),
                ("x_dim", C.c_int32), ("u_dim", C.c_int32), ("N", C.c_int32)]


class ProjSetEntry(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rows", C.c_int32), ("A", C.c_void_p), ("b", C.c_void_p), ("p0", C.c_void_p),
                ("p1", C.c_void_p), ("l", C.c_double), ("u", C.c_double)]


class ProjSetParams(_Guarded):
    _fields_ = [("struct_size", C.c_uint32), ("n_sets", C.c_int32), ("sets", ProjSetEntry * 4), ("rho", C.c_double),
                ("threshold", C.c_double), ("max_iter", C.c_int32)]


class SolveOut(_Guarded):
    _fields_ = [("struct_size", C.c_uint32)] + [(f, C.c_void_p) for f in OUT_FIELDS]


EXPORTS = ["isls_version", "isls_last_error_string", "isls_model_id", "isls_model_supported", "isls_plan_create",
           "isls_plan_destroy", "isls_workspace_bytes", "isls_ilqr_admm_solve_f64", "isls_ilqr_solve_f64",
           "isls_lqt_admm_dp_f64", "isls_riccati_f64", "isls_rollout_linesearch_f64", "isls_admm_project_dual_f64",
           "isls_measure_fp64_tflops", "isls_profile_enable", "isls_profile_collect",
           "isls_sls_plan_create", "isls_sls_plan_destroy", "isls_sls_operators", "isls_sls_solve_f64",
           "isls_sls_admm_f64", "isls_sls_controller_f64", "isls_mc_rollout_f64", "isls_project_rows_f64",
           "isls_isls_admm_solve_f64", "isls_sls_replan_f64", "isls_probe_overlap_f64", "isls_project_rows_ex_f64", "isls_project_set_convex_f64",
           "isls_controller_tv_workspace_bytes", "isls_controller_tv_f64", "isls_linearize_f64"]

KERNEL_CLASSES = ["init", "kpass", "ff", "linesearch", "admm", "outer_end", "finalize", "backward_full", "accept",
                  "lqt", "compact", "isls_cols", "isls_update", "admm_loop"]

_lib = None


class IslsError(RuntimeError):
    pass


def lib():
    """Load libisls_b200.so (once).  Raises if it has not been built - there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IslsError("libisls_b200.so not found at %s - run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.isls_last_error_string.restype = C.c_char_p
    L.isls_model_id.argtypes = [C.c_char_p]
    L.isls_model_supported.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.isls_plan_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(C.c_void_p)]
    L.isls_plan_destroy.argtypes = [C.c_void_p]
    L.isls_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]
    solve_args = [C.c_void_p, C.POINTER(SolveOpts), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_size_t, C.POINTER(SolveOut), C.c_void_p]
    L.isls_ilqr_admm_solve_f64.argtypes = solve_args
    L.isls_isls_admm_solve_f64.argtypes = [C.c_void_p, C.POINTER(SolveOpts), C.POINTER(SlsAdmmOpts), C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(SolveOut),
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.isls_ilqr_solve_f64.argtypes = solve_args
    L.isls_probe_overlap_f64.argtypes = solve_args[:9] + [C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_void_p]
    L.isls_lqt_admm_dp_f64.argtypes = [C.c_void_p, C.POINTER(SolveOpts), C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.POINTER(SolveOut), C.c_void_p]
    L.isls_riccati_f64.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 8
    L.isls_rollout_linesearch_f64.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 11 + [C.c_size_t, C.c_void_p]
    L.isls_admm_project_dual_f64.argtypes = [C.c_int64, C.c_int64, C.c_double] + [C.c_void_p] * 9
    L.isls_measure_fp64_tflops.argtypes = [C.POINTER(C.c_double), C.c_void_p]
    L.isls_sls_plan_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_double, C.POINTER(C.c_void_p), C.c_void_p]
    L.isls_sls_plan_destroy.argtypes = [C.c_void_p]
    L.isls_sls_operators.argtypes = [C.c_void_p] * 5
    L.isls_sls_solve_f64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.isls_sls_admm_f64.argtypes = [C.c_void_p, C.POINTER(SlsAdmmOpts), C.c_int64] + [C.c_void_p] * 8
    L.isls_sls_controller_f64.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.isls_sls_replan_f64.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    L.isls_controller_tv_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64]
    L.isls_controller_tv_workspace_bytes.restype = C.c_size_t
    L.isls_controller_tv_f64.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    L.isls_linearize_f64.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    L.isls_mc_rollout_f64.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                      C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.isls_project_rows_f64.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_double, C.c_void_p, C.c_void_p]
    L.isls_project_rows_ex_f64.argtypes = [C.POINTER(ProjParams), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
    L.isls_project_set_convex_f64.argtypes = [C.POINTER(ProjSetParams), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
    L.isls_profile_enable.argtypes = [C.c_int]
    L.isls_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().isls_last_error_string()
        raise IslsError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
