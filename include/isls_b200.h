/* isls_b200.h - C-ABI of the B200-native batched iLQR-ADMM hot path (libisls_b200.so).
 *
 * The reference (chenjianxing1/iLQR-ADMM, package `isls`) is pure Python: its "FFI" for this path is the Python
 * class surface.  Each entry point below names the reference interface it replaces (file:line in the
 * reference tree).  The host-side mirror of that surface (ilqr-admm_b200/isls_b200: iSLS, SLS, ADMM) binds these
 * functions with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no C++/torch types.  All numeric arrays are IEEE FP64 (`double`) or `int32_t`.
 *   - "dev" pointers are DEVICE pointers owned by the caller; "host" pointers are small host arrays that are
 *     copied when a plan is created.  Natural (reference) layouts: row-major [B, N, dim] etc.
 *   - the library never allocates device memory on the solve path: the caller provides a workspace of
 *     isls_workspace_bytes() bytes (256-byte aligned).  Plans own a small constant block (created / destroyed
 *     explicitly).
 *   - every solve call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and returns.
 *   - return value: 0 ok; <0 invalid argument / unsupported (see ISLS_E_*); >0 a cudaError_t.
 *     isls_last_error_string() describes the last failure on the calling thread.
 *   - per-problem conditions are reported in `status[B]` bit-fields (ISLS_ST_*), never as return codes.
 *   - ABI guard: every POD struct starts with `uint32_t struct_size`, which the caller sets to sizeof(the struct) as
 *     declared in the header it was built against.  A mismatch (stale binding) is rejected with ISLS_E_INVALID
 *     instead of reading past the caller's struct.  Use ISLS_INIT(type, var) in C.
 */
#ifndef ISLS_B200_H
#define ISLS_B200_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISLS_VERSION 200

/* return codes */
#define ISLS_OK 0
#define ISLS_E_INVALID (-1)      /* bad argument (NULL pointer, non-positive size, ...) */
#define ISLS_E_UNSUPPORTED (-2)  /* unknown model / unsupported (n, m) */
#define ISLS_E_WORKSPACE (-3)    /* workspace too small or misaligned */

/* zero-initialised struct with its ABI guard set:  ISLS_INIT(isls_solve_opts, o); o.max_outer = 20; ... */
#define ISLS_INIT(type, var) type var; memset(&(var), 0, sizeof(var)); (var).struct_size = (uint32_t)sizeof(var)

/* device-side dynamics models, registered by name behind the reference's forward_model / get_AB plugin slots
 * (isls/isls_base.py:106-111 forward_model; get_AB argument of isls/isls.py:54, isls/isls.py:379) */
#define ISLS_MODEL_DOUBLE_INTEGRATOR 0   /* isls/utils.py:266-276 + isls/sls_base.py:49-53; n=2d, m=d, d in {1,2,3} */
#define ISLS_MODEL_CAR 1                 /* notebooks/Car/Iterative LQR with control constraints.ipynb cell 6; n=4,m=2 */
#define ISLS_MODEL_ARM3 2                /* notebooks/3DoF robot/State and control bound constraints.ipynb cells 9-10; n=9,m=3 */
#define ISLS_MODEL_TASSA_CAR 3           /* notebooks/Tutorial.ipynb cell 8 (car parking of Tassa et al., axle distance 2); n=4,m=2 */
#define ISLS_MODEL_LTI 4                 /* x+ = A x + B u with any constant A, B (Base.AB, isls/base.py:98-119; name "lti");
                                            (n,m) in {(2,1),(4,2),(6,3)}; matrices in isls_problem_desc.lti_A / lti_B */

/* state-cost families (cost_function slot, isls/isls_base.py:113-131) */
#define ISLS_COST_QUADRATIC 0            /* sum_i Q_ii (x_i - z_i)^2: isls/sls_base.py:25-44 */
#define ISLS_COST_PSEUDO_HUBER 1         /* sum_i w_i (sqrt((x_i - z_i)^2 + p_i^2) - p_i), two weighted terms per
                                            component: notebooks/Tutorial.ipynb cell 14 (running + final cost) */

/* per-problem status bits */
#define ISLS_ST_CONVERGED_COST 1   /* |cost-prev| < tol            isls/isls.py:125, isls/isls.py:493 */
#define ISLS_ST_LINESEARCH_FAIL 2  /* forward pass failed           isls/isls.py:128 */
#define ISLS_ST_MAX_ITER 4         /* iteration budget exhausted    isls/isls.py:131 */
#define ISLS_ST_OSCILLATING 8      /* mean-of-4 oscillation test    isls/isls.py:497 */
#define ISLS_ST_NON_PD 16          /* Quu not positive definite (the reference raises LinAlgError, isls/isls.py:296) */
#define ISLS_ST_NAN_COST 32        /* NaN cost in the line search   isls/isls.py:362 */

/* ADMM exit codes (admm_exit[B, max_outer]) */
#define ISLS_ADMM_CONVERGED 1      /* isls/admm.py:72 */
#define ISLS_ADMM_STALLED 2        /* isls/admm.py:80 */
#define ISLS_ADMM_MAXIT 3          /* isls/admm.py:93 */

/* Problem description shared by all problems of a batch: dynamics model, horizon, quadratic via-point cost
 * (isls/base.py:81-89 set_quadratic_cost), diagonal ADMM penalties (isls/base.py:55-79 compute_Rr_Qr, diagonal
 * rho only - see SURVEY D10) and box bounds (isls/projections.py:7-11 project_bound).  All pointers are HOST
 * pointers, copied by isls_plan_create. */
typedef struct isls_problem_desc {
  uint32_t struct_size;  /* = sizeof(isls_problem_desc) (ABI guard) */
  int32_t model_id;      /* ISLS_MODEL_* */
  int32_t n, m, N;       /* x_dim, u_dim, horizon (N states x_0..x_{N-1}, N controls) */
  int32_t n_via;         /* number of via-points k (rows of zs / Qdiag) */
  int32_t L;             /* number of line-search candidates (<= 50, isls/isls_base.py:10-11) */
  double dt;             /* model time step */
  double u_std;          /* R = u_std * I (isls/base.py:86) */
  const double *Qdiag;   /* [n_via, n] diagonals of Qs */
  const int32_t *seq;    /* [N] via-point index per time step */
  const double *alphas;  /* [L] step sizes, 10**linspace(0,-5,50)[:L] */
  const double *rho_x;   /* [N, n] diagonal of Qr per step, or NULL: no state projection */
  const double *lo_x, *hi_x; /* [N, n] box bounds on the state (+-inf = free), required iff rho_x != NULL */
  const double *rho_u;   /* [N, m] or NULL: no control projection */
  const double *lo_u, *hi_u; /* [N, m] */
  /* ---- optional cost extensions (all 0 / NULL: quadratic via-point cost with R = u_std I) ---- */
  int32_t cost_kind;     /* ISLS_COST_*: for PSEUDO_HUBER Qdiag holds the weights w of term a */
  const double *Rdiag;   /* [m] diagonal of R (overrides u_std; Tutorial cell 14: cu) or NULL */
  const double *Hp;      /* [n_via, n] smoothness scales p of term a (PSEUDO_HUBER) */
  const double *Qdiag_b; /* [n_via, n] weights of the second term per component or NULL */
  const double *Hp_b;    /* [n_via, n] smoothness scales of the second term */
  /* ---- optional state projection onto the outside of obstacle sets instead of the box lo_x / hi_x (rho_x still
   * required): project_set_convex (isls/projections.py:289-374) with As = I, bs = 0 over n_obst sets, each the
   * infinity-norm shell lower_k <= ||W_k (p - c_k)||_inf <= upper of the position p = x[:2]
   * (project_square_batch, isls/projections.py:246-255; notebooks/Car/Iterative LQR with state constraints.ipynb
   * cell 18).  All rows (time steps) of a problem are projected together: the inner ADMM stops on the maximum of the
   * residual norms over sets and rows. ---- */
  int32_t n_obst;              /* 0: box projection; 1..ISLS_MAX_OBST */
  int32_t obst_max_iter;       /* max_iter of the inner ADMM */
  const double *obst_centers;  /* [n_obst, 2] */
  const double *obst_W;        /* [n_obst, 2, 2] row-major */
  const double *obst_W_inv;    /* [n_obst, 2, 2] */
  const double *obst_lower;    /* [n_obst] */
  double obst_upper, obst_rho, obst_threshold;
  int32_t obst_kind;           /* 0: rotated infinity-norm shells on all n components (As = I_n, car parking notebook);
                                  1: quadratic shells obst_lower <= 0.5 ||p - c||^2 <= obst_upper of the position only
                                  (As = I_2; project_quadratic, isls/projections.py:91-105; obst_W unused), followed by
                                  project_set_convex_dykstra (isls/projections.py:465-505) when obst_dykstra_max_iter > 0
                                  (Double integrator/LQR and SLS with spherical obstacle avoidance.ipynb cell 12;
                                  implemented on the LQT path, isls_lqt_admm_dp_f64) */
  int32_t obst_dykstra_max_iter;
  double obst_dykstra_tol;
  int32_t isls_dim;            /* > 0: the plan is used by isls_isls_admm_solve_f64 with `dim` robustness columns
                                  (workspace for the [d_u | Phi_u(:, :dim)] matrix variables); <= 3 */
  const double *lti_A, *lti_B; /* ISLS_MODEL_LTI: A [n, n], B [n, m] row-major (host; copied) */
} isls_problem_desc;
#define ISLS_MAX_OBST 4

typedef struct isls_plan isls_plan;   /* opaque */

/* iteration budgets and tolerances: keyword arguments of iSLS.ilqr_admm (isls/isls.py:379-381) and
 * iSLS.solve (isls/isls.py:54-55) */
typedef struct isls_solve_opts {
  uint32_t struct_size;    /* = sizeof(isls_solve_opts) (ABI guard) */
  int32_t max_outer;       /* max_iter */
  int32_t max_admm;        /* max_admm_iter (ignored by isls_ilqr_solve_f64) */
  double tol;              /* ADMM residual tolerance `tol` (admm.py:72-85) / tol_fun for plain iLQR */
  double outer_tol;        /* hard-coded 1e-3 in isls/isls.py:493,497 */
  double relax;            /* ADMM relaxation alpha (admm.py:46) */
  int32_t fixed_budget;    /* 1: ignore every stop test (deterministic work, used by the benchmark) */
  int32_t last_stage_dp;   /* 0: batch-form last control du_{N-1} = -Cuu^-1 cu (isls/isls.py:441-465);
                              1: DP form K[N-1]=k[N-1]=0 (isls/isls.py:245-246) */
  double stall_tol;        /* ADMM relative-change stop threshold; 0: = tol (admm.py:80); isls_admm: 1e-3 (isls.py:664) */
  double osc_tol;          /* oscillation test threshold; 0: = outer_tol; isls_admm: 1e-3 with outer_tol 1e-4 */
  const double *z_x_init_dev; /* isls_lqt_admm_dp_f64 only: ADMM warm start z_x [B, N, n] / z_u [B, N, m] (DEVICE pointers, */
  const double *z_u_init_dev; /* natural layout) or NULL = zeros.  ADMM_LQT_Batch starts from the unconstrained solution
                                 (isls/sls.py:266-268) and uses last_stage_dp = 0. */
} isls_solve_opts;

/* Results, natural layouts, DEVICE pointers.  Optional outputs may be NULL. */
typedef struct isls_solve_out {
  uint32_t struct_size; /* = sizeof(isls_solve_out) (ABI guard) */
  double *x;            /* [B, N, n] final nominal states            (iSLS.x_nom) */
  double *u;            /* [B, N, m] final nominal controls          (iSLS.u_nom) */
  double *cost;         /* [B] final cost                            (iSLS.cost) */
  double *cost_log;     /* [B, max_outer+1], NaN padded              (iSLS.cost_log) */
  int32_t *n_log;       /* [B] valid entries of cost_log */
  int32_t *status;      /* [B] ISLS_ST_* */
  int32_t *outer_iters; /* [B] */
  int32_t *admm_iters;  /* [B, max_outer] optional */
  int32_t *admm_exit;   /* [B, max_outer] optional */
  double *res_log;      /* [B, max_outer, max_admm, 2] optional: (primal, dual) residual norms (ADMM `logs`) */
  int32_t *alpha_idx;   /* [B, max_outer, max_admm] optional: chosen line-search index (-1: rejected) */
  double *z_x, *z_u;    /* [B, N, n], [B, N, m] optional: final ADMM z (warm-start state, isls.py:489-490) */
  double *lam_x, *lam_u;/* optional: final scaled duals */
  double *K;            /* [B, N, m, n] optional: feedback gains of the last backward pass */
  double *k;            /* [B, N, m] optional: feed-forward gains of the last backward pass */
  int8_t *mask_x, *mask_u; /* [B, N, dim] optional: clip mask of the last projection (-1 at lo, +1 at hi) */
  int32_t *inner_iters; /* [B, max_outer, max_admm] optional: iterations of the inner project_set_convex ADMM
                           (obstacle-set state projection only) */
  double *Quu, *Quu_inv;/* [B, N, m, m] optional: Quu_t and its inverse of the last backward pass (the `return_Qs` logs of */
  double *Qux;          /* SLS.solve_dp, isls/sls.py:117-120, 159-162); Qux [B, N, m, n].  Row N-1 is zero. */
} isls_solve_out;

int isls_version(void);
const char *isls_last_error_string(void);

/* name -> ISLS_MODEL_* (or ISLS_E_UNSUPPORTED).  Names: "double_integrator", "car", "arm3", "tassa_car", "lti". */
int isls_model_id(const char *name);
/* 0 if the (model, n, m) combination has a compiled kernel */
int isls_model_supported(int32_t model_id, int32_t n, int32_t m);

int isls_plan_create(const isls_problem_desc *desc, isls_plan **plan);
int isls_plan_destroy(isls_plan *plan);

/* bytes of caller-owned device workspace needed to solve B problems with this plan */
int isls_workspace_bytes(const isls_plan *plan, int64_t B, size_t *bytes);

/* iSLS.ilqr_admm (isls/isls.py:379-501) with ADMM (isls/admm.py:6-106) and project_bound
 * (isls/projections.py:7-11) for a batch of B independent problems.
 *   x0_dev [B, n]; u_init_dev [B, N, m]; zs_dev [B, n_via, n] via-point targets per problem. */
int isls_ilqr_admm_solve_f64(const isls_plan *plan, const isls_solve_opts *opts, int64_t B,
                             const double *x0_dev, const double *u_init_dev, const double *zs_dev,
                             void *workspace_dev, size_t workspace_bytes, const isls_solve_out *out, void *stream);

/* Method isls_admm of iSLS, isls/isls.py:503-712: robust nonlinear iSLS-ADMM on the matrix variable [d_u | Phi_u(:, :dim)]
 * (robustness with respect to the first `dim` components of the initial state) with the row-wise projection
 * project_u(z, u_nom) = project_set_convex over second-order cones (3DoF robot/State bounds and robust control
 * bounds.ipynb cells 24-26; `soc` describes the cones like isls_sls_admm_opts: n_cones, cone_rows = dim + 2, As, bs,
 * inner_rho, inner_max_iter, inner_threshold; its other fields are ignored).  Riccati form: one K-pass per outer
 * iteration, dim + 1 feed-forward passes + linear rollouts per ADMM iteration instead of the dense (N m)^2 inverse of
 * isls.py:562-579.  The plan needs rho_u (Rr = diag) and isls_dim = dim; N * m <= 1024.
 * State side (project_x, isls.py:556-559, 571-572, 631-638, 648-650): plan with rho_x (Qr = diag, zero on the rows that
 * are not penalised) and soc->n_x_rows > 0, read here as: x_row_idx[g] = a projected state COMPONENT (<= 8 of them),
 * x_bs [n_x_rows, P, cone_rows] = the cone offsets of that component (the A_i are shared with the control side),
 * rho_x_rows not read.  project_x(z, x_nom) shifts column 0 by x_nom, projects the N rows of every listed component of
 * [d_x | Phi_x(:, :dim)] with one project_set_convex call per component and passes all other rows through.
 * soc->bs == NULL with n_cones > 0: the control side is not projected (the isls_admm method called with project_x alone).
 *   du_dev [B, N, m] (= x_u[:, 0]), phi_u_dev [B, N, m, dim] (= x_u[:, 1:]) of the last ADMM iterate. */
struct isls_sls_admm_opts;
int isls_isls_admm_solve_f64(const isls_plan *plan, const isls_solve_opts *opts, const struct isls_sls_admm_opts *soc,
                             int64_t B, const double *x0_dev, const double *u_init_dev, const double *zs_dev,
                             void *workspace_dev, size_t workspace_bytes, const isls_solve_out *out, double *du_dev,
                             double *phi_u_dev, void *stream);

/* iSLS.solve(method='dp') (isls/isls.py:54-132, 229-374): unconstrained iLQR with closed-loop line search. */
int isls_ilqr_solve_f64(const isls_plan *plan, const isls_solve_opts *opts, int64_t B,
                        const double *x0_dev, const double *u_init_dev, const double *zs_dev,
                        void *workspace_dev, size_t workspace_bytes, const isls_solve_out *out, void *stream);

/* SLS.ADMM_LQT_DP (isls/sls.py:298-317; solve_dp 85-166, solve_dp_ff 168-202): linear dynamics
 * (double integrator), opts->max_admm iterations, single "outer" pass.  out->cost_log gets [B,2]. */
int isls_lqt_admm_dp_f64(const isls_plan *plan, const isls_solve_opts *opts, int64_t B,
                         const double *x0_dev, const double *zs_dev,
                         void *workspace_dev, size_t workspace_bytes, const isls_solve_out *out, void *stream);

/* ---- stage-level entry points (teacher-forced parity tests; also usable on their own) ---- */

/* iSLS.backward_pass_DP(Cts, cts) (isls/isls.py:229-308), generic-operator variant: dense operators from HBM.
 *   A_dev [B,N,n,n], B_dev [B,N,n,m], c_dev [B,N,n+m], C_dev [B,N,n+m,n+m] -> K_dev [B,N,m,n], k_dev [B,N,m],
 *   non_pd_dev [B] (int32, optional).  Supported (n,m): (2,1) (4,2) (6,3) (9,3). */
int isls_riccati_f64(int32_t n, int32_t m, int32_t N, int64_t B, const double *A_dev, const double *B_dev,
                     const double *c_dev, const double *C_dev, double *K_dev, double *k_dev, int32_t *non_pd_dev,
                     void *stream);

/* iSLS.rollout_batch + cost evaluation + argmin (isls/isls.py:135-154, 468-477): open-loop line search from
 * x_nom[:,0] over candidates u_nom + alpha_l * du with the ADMM penalty terms.
 *   x_nom_dev [B,N,n] (only x_nom[:,0] is used, like the reference), u_nom_dev [B,N,m], du_dev [B,N,m],
 *   reg_x_dev [B,N,n] / reg_u_dev [B,N,m] (z - lambda; may be NULL when the plan has no such projection)
 *   -> costs_dev [B,L], best_dev [B] (int32), x_best_dev [B,N,n], u_best_dev [B,N,m]. */
int isls_rollout_linesearch_f64(const isls_plan *plan, int64_t B, const double *x_nom_dev, const double *u_nom_dev,
                                const double *du_dev, const double *zs_dev, const double *reg_x_dev,
                                const double *reg_u_dev, double *costs_dev, int32_t *best_dev, double *x_best_dev,
                                double *u_best_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* One ADMM z-projection + scaled-dual update (isls/admm.py:43-69) with project_bound, elementwise on flat
 * [B, len] arrays with per-element bounds lo/hi [len] (device):
 *   z <- clip(relax*x + (1-relax)*z + lam, lo, hi); lam <- lam + x - z;
 *   prim_sq[B] += sum (x-z)^2, dual_sq[B] += sum (z-z_prev)^2, mask[B,len] (optional) = -1/0/+1.
 * Bit-exact with numpy on the same inputs (no FMA contraction). */
int isls_admm_project_dual_f64(int64_t B, int64_t len, double relax, const double *x_dev, double *z_dev,
                               double *lam_dev, const double *lo_dev, const double *hi_dev, double *prim_sq_dev,
                               double *dual_sq_dev, int8_t *mask_dev, void *stream);

/* ---- SLS path (north-star item 4, BASELINE config 4): linear time-invariant dynamics, diagonal Q ---- */
typedef struct isls_sls_plan isls_sls_plan;   /* opaque: Sw, Su, Su'Q, L = Su'QSu + R, L^-1, PHI_U on the device */

/* Builds the operators shared by a batch (same A, B, Q, R; per-problem targets xd):
 *   Sw, Su                         Base.AB setter            isls/base.py:98-119
 *   L = Su'Q Su + R, L^-1          SLS.solve_sls             isls/sls.py:216-219 (compute_inverses, base.py:44-50)
 *   PHI_U (block lower triangular) SLS.solve_sls             isls/sls.py:225-229
 * A_host [n,n], B_host [n,m], Qdiag_t_host [N,n] (per-step diagonal of Q) are HOST arrays.  Synchronises. */
int isls_sls_plan_create(int32_t n, int32_t m, int32_t N, const double *A_host, const double *B_host,
                         const double *Qdiag_t_host, double u_std, isls_sls_plan **plan, void *stream);
int isls_sls_plan_destroy(isls_sls_plan *plan);
/* copies the shared operators into caller-owned device buffers (any may be NULL):
 * Sw [N n, N n], Su [N n, N m], PHI_U [N m, N n], row-major */
int isls_sls_operators(const isls_sls_plan *plan, double *Sw_dev, double *Su_dev, double *PHI_U_dev, void *stream);

/* SLS.solve_sls feed-forward part (isls/sls.py:221): du[b] = L^-1 Su'Q xd[b].  xd_dev [B, N n] -> du_dev [B, N m] */
int isls_sls_solve_f64(const isls_sls_plan *plan, int64_t B, const double *xd_dev, double *du_dev, void *stream);

/* SLS.ADMM_SLS (isls/sls.py:319-454) with project_u = row-wise project_set_convex(.., [project_soc_unit]*P)
 * (isls/projections.py:289-374, 140-162) and no state projection: robust control bounds w.r.t. the initial position. */
typedef struct isls_sls_admm_opts {
  uint32_t struct_size;    /* = sizeof(isls_sls_admm_opts) (ABI guard) */
  int32_t max_iter;        /* ADMM_SLS max_iter */
  double rho_u, alpha, tol;
  int32_t fixed_budget;    /* 1: ignore the stop tests */
  int32_t n_cones;         /* P: number of cones per row (<= 4) */
  int32_t cone_rows;       /* rows of each A_i = c + 1 with c = 1 + x_dim/2 columns (c <= 4) */
  const double *As;        /* host [P, cone_rows, c] */
  const double *bs;        /* host [P, cone_rows] */
  double inner_rho;        /* project_set_convex rho */
  int32_t inner_max_iter;  /* project_set_convex max_iter */
  double inner_threshold;  /* project_set_convex threshold */
  /* ---- optional state-side projection project_x (sls.py:342-347, 381-382, 396-401, 414-415; Double integrator/LQR
   * and SLS with state bounds.ipynb cells 16-17): the listed rows of [d_x | Phi_x(:, :n/2)] are each projected by
   * their own project_set_convex call onto {A_i x + b_i in SOC} (same A_i as the control side, own b_i); Qr = diag
   * with rho_x_rows on those rows and 0 elsewhere.  n_x_rows = 0: no state projection. ---- */
  int32_t n_x_rows;          /* <= ISLS_MAX_XROWS */
  const int32_t *x_row_idx;  /* host [n_x_rows]: row index into N * n */
  const double *x_bs;        /* host [n_x_rows, P, cone_rows] */
  const double *rho_x_rows;  /* host [n_x_rows]: Qr diagonal entries of those rows */
} isls_sls_admm_opts;
#define ISLS_MAX_XROWS 8
/* xd_dev [B, N n] -> du_dev [B, N m] (= x_u[:,0]), phi_cols_dev [B, N m, c-1] (= x_u[:,1:c]; the full PHI_U of a
 * problem is [phi_cols | shared PHI_U[:, c-1:]], sls.py:450), logs_dev [B, max_iter, 2] (optional), iters_dev [B],
 * exit_dev [B] (ISLS_ADMM_*), inner_total_dev [B] int64 (optional: total inner projection iterations). */
int isls_sls_admm_f64(isls_sls_plan *plan, const isls_sls_admm_opts *opts, int64_t B, const double *xd_dev,
                      double *du_dev, double *phi_cols_dev, double *logs_dev, int32_t *iters_dev, int32_t *exit_dev,
                      int64_t *inner_total_dev, void *stream);

/* SLS.controller (isls/sls.py:235-242): K = PHI_U PHI_X^-1, k = (I - K Su) du per problem, with
 * PHI_U[b] = [phi_cols_dev[b] (first n_first_cols columns) | shared PHI_U (rest)].
 * workspace: B * (N n)^2 doubles.  K_dev [B, N m, N n], k_dev [B, N m]. */
int isls_sls_controller_f64(const isls_sls_plan *plan, int64_t B, int32_t n_first_cols, const double *phi_cols_dev,
                            const double *du_dev, void *workspace_dev, size_t workspace_bytes, double *K_dev,
                            double *k_dev, void *stream);

/* iSLS.controller for a general causal PHI_U on time-varying dynamics: the same formula (isls/sls.py:235-242) with the
 * operators C = (I - Z A_d)^-1, D = C Z B_d that iSLSBase.AB builds from A_t, B_t (isls/isls_base.py:138-158).
 *   A_dev [B, N, n, n], B_dev [B, N, n, m] (or [N, n, n], [N, n, m] shared by the batch when shared_AB != 0; entry N-1
 *   is not read), PHI_U_dev [B, N m, N n] (block lower triangular: u_t reacts to w_s for s <= t), du_dev [B, N m]
 *   -> K_dev [B, N m, N n], k_dev [B, N m].  workspace: isls_controller_tv_workspace_bytes(). */
size_t isls_controller_tv_workspace_bytes(int32_t n, int32_t m, int32_t N, int64_t B);
int isls_controller_tv_f64(int32_t n, int32_t m, int32_t N, int64_t B, const double *A_dev, const double *B_dev,
                           int32_t shared_AB, const double *PHI_U_dev, const double *du_dev, void *workspace_dev,
                           size_t workspace_bytes, double *K_dev, double *k_dev, void *stream);

/* get_AB of the reference's callers on the device: A_t = df/dx, B_t = df/du of a registered model at every row of
 * x_dev [rows, n], u_dev [rows, m] -> A_dev [rows, n, n], B_dev [rows, n, m] (the layouts iSLSBase.AB takes,
 * isls/isls_base.py:133-158). */
int isls_linearize_f64(int32_t model_id, int32_t n, int32_t m, double dt, int64_t rows, const double *x_dev,
                       const double *u_dev, double *A_dev, double *B_dev, void *stream);

/* SLS.initialize_replanning_procedure + SLS.replan_feedforward (isls/sls.py:244-248): new feed-forward terms for a new
 * target vector without re-solving,  k_new = k + (I - K Su) (Su'Q Su + R)^-1 Su'Q (xd_new - xd_old),  evaluated as four
 * matrix-vector products per problem (the replan matrix is never formed).
 *   K_dev [B, N m, N n], k_dev [B, N m], xd_new_dev / xd_old_dev [B, N n] -> k_new_dev [B, N m]. */
int isls_sls_replan_f64(const isls_sls_plan *plan, int64_t B, const double *K_dev, const double *k_dev,
                        const double *xd_new_dev, const double *xd_old_dev, double *k_new_dev, void *stream);

/* ---- Monte-Carlo closed-loop evaluation of one controller over B sampled initial states ----
 * mode 0: get_trajectory_batch  u_t = us[t]                         (k_dev = us [N, m])
 * mode 1: get_trajectory_dp     u_t = K_t x_t + k_t                 (K_dev [N, m, n], k_dev [N, m])
 * mode 2: get_trajectory_sls    u_t = sum_s K[t,s](x_s - x^_s) + k_t + u^_t   (K_dev [N m, N n], k_dev [N m];
 *                                x_nom_dev [N, n] / u_nom_dev [N, m] optional: iSLSBase subtracts the nominal)
 * isls/sls_base.py:62-105, isls/isls_base.py:28-71.  x_{t+1} = f(x_t, u_t) + w, w ~ N(0, noise_scale) from a
 * counter-based generator keyed by (seed, sample, step).  x0_dev [B, n] -> x_out_dev [B, N, n], u_out_dev [B, N, m]. */
int isls_mc_rollout_f64(int32_t model_id, int32_t n, int32_t m, int32_t N, double dt, int32_t mode, int64_t B,
                        const double *x0_dev, const double *K_dev, const double *k_dev, const double *x_nom_dev,
                        const double *u_nom_dev, double noise_scale, uint64_t seed, double *x_out_dev,
                        double *u_out_dev, void *stream);

/* ---- batched row projections: device counterparts of the `_batch` functions of isls/projections.py (8f #2) ----
 * x_dev [rows, dim] -> out_dev [rows, dim] (dim <= 16), one row per thread.
 *   kind 0 bound      np.clip(x, p0 = lo[dim], p1 = hi[dim])                     isls/projections.py:7-11
 *   kind 1 linear     l <= a'x <= u, a = p0[dim]                                  isls/projections.py:30-43
 *   kind 2 quadratic  l <= 0.5 |x - c|^2 <= u, c = p0[dim] or NULL                isls/projections.py:86-104
 *   kind 3 soc_unit   |x[:-1]| <= x[-1], numpy batch semantics                    isls/projections.py:140-162
 *   kind 4 square     l <= |x - c|_inf <= u, c = p0[dim] or NULL                  isls/projections.py:252-272
 *   kind 5 unit_ball  |x| <= 1                                                    isls/projections.py:232-240 */
int isls_project_rows_f64(int32_t kind, int64_t rows, int32_t dim, const double *x_dev, const double *p0_dev,
                          const double *p1_dev, double l, double u, double *out_dev, void *stream);

/* Parameterised row projections (8f #2, continued).  All pointers are HOST arrays (small parameters, copied). */
#define ISLS_PROJ_MULTILINEAR 6              /* l <= A x <= u, boundary projection x - A'(AA')^-1 (Ax - clip)   isls/projections.py:46-62 */
#define ISLS_PROJ_SOC 7                      /* A x + b in SOC by the inner ADMM of project_soc (rows <= 1024: its stop
                                                rule is a maximum over all rows)                                isls/projections.py:163-232 */
#define ISLS_PROJ_BLOCK_LOWER_TRIANGULAR 8   /* z[i*u_dim, i*x_dim:(i+1)*x_dim] = 0 in place on out_dev [N u_dim, N x_dim]
                                                                                                                isls/projections.py:277-286 */
typedef struct isls_proj_params {
  uint32_t struct_size;  /* = sizeof(isls_proj_params) (ABI guard) */
  int32_t kind;          /* ISLS_PROJ_* */
  int32_t k;             /* rows of A (<= 8) */
  const double *A;       /* [k, dim] */
  const double *b;       /* [k] (soc) or NULL */
  const double *l, *u;   /* [k] (multilinear) or NULL = unbounded */
  double rho, tol;       /* soc: inner ADMM penalty and tolerance */
  int32_t max_iter;      /* soc */
  int32_t x_dim, u_dim, N;   /* block_lower_triangular */
} isls_proj_params;
/* x_dev [rows, dim] -> out_dev [rows, dim] (dim <= 16); iters_dev [1] optional (soc: inner iterations).
 * project_affine (isls/projections.py:64-68) is kind 1 of isls_project_rows_f64 with shifted bounds l - b, u - b. */
int isls_project_rows_ex_f64(const isls_proj_params *params, int64_t rows, int32_t dim, const double *x_dev,
                             double *out_dev, int32_t *iters_dev, void *stream);

/* Generic project_set_convex (isls/projections.py:289-374): rows x0 onto the intersection of up to 4 sets
 * {x : A_i x + b_i in C_i} by the reference's consensus ADMM, every C_i one of the primitive row projections of
 * isls_project_rows_f64 (kind 0 bound with p0 = lo, p1 = hi [rows_i]; 2 quadratic shell l <= 0.5|y - p0|^2 <= u; 3 soc_unit;
 * 4 infinity-norm shell l <= |y - p0|_inf <= u; 5 unit ball), batch semantics.  The stop rule is a maximum over sets and
 * rows, so all rows (<= 1024) are projected together.  All pointers are HOST arrays. */
typedef struct isls_proj_set_entry {
  int32_t kind;          /* primitive projection of this set (see above) */
  int32_t rows;          /* rows of A_i (<= 8) */
  const double *A;       /* [rows, dim] */
  const double *b;       /* [rows] or NULL = 0 */
  const double *p0, *p1; /* [rows] parameters of the primitive, or NULL */
  double l, u;
} isls_proj_set_entry;
typedef struct isls_proj_set_params {
  uint32_t struct_size;  /* = sizeof(isls_proj_set_params) (ABI guard) */
  int32_t n_sets;        /* 1..4 */
  isls_proj_set_entry sets[4];
  double rho, threshold;
  int32_t max_iter;
} isls_proj_set_params;
int isls_project_set_convex_f64(const isls_proj_set_params *params, int64_t rows, int32_t dim, const double *x_dev,
                                double *out_dev, int32_t *iters_dev, void *stream);

/* ---- measurement helpers (bench.py roofline denominators; not part of the reference surface) ---- */
/* kernel classes for per-kernel CUDA-event timing */
#define ISLS_KC_INIT 0
#define ISLS_KC_KPASS 1
#define ISLS_KC_FF 2
#define ISLS_KC_LINESEARCH 3
#define ISLS_KC_ADMM 4
#define ISLS_KC_OUTER_END 5
#define ISLS_KC_FINALIZE 6
#define ISLS_KC_BACKWARD_FULL 7
#define ISLS_KC_ACCEPT 8
#define ISLS_KC_LQT 9
#define ISLS_KC_COMPACT 10
#define ISLS_KC_ISLS_COLS 11
#define ISLS_KC_ISLS_UPDATE 12
#define ISLS_KC_ADMM_LOOP 13
#define ISLS_KC_COUNT 14
/* thread-local switch: when on, every kernel launch of a solve is bracketed by a CUDA event pair on the
 * launching stream (adds a few microseconds per launch; use for per-kernel durations, not for throughput) */
int isls_profile_enable(int on);
/* synchronises, sums elapsed milliseconds and launch counts per kernel class ([ISLS_KC_COUNT] each), resets */
int isls_profile_collect(double *ms_sum, int64_t *launches);
/* durations (ms, best of 3; synchronous) of the two kernels of one slot of the overlapped large-batch schedule, each alone
 * on its half of the tiles and both together: ms_host[6] = line search (CTA per tile), persistent line search with
 * ls_ctas CTAs per SM, TMA-staged ff-pass with ring depth ff_depth, plain ff-pass, persistent line search || TMA
 * ff-pass, plain line search || TMA ff-pass.  Arguments as isls_ilqr_admm_solve_f64. */
int isls_probe_overlap_f64(const isls_plan *plan, const isls_solve_opts *opts, int64_t B, const double *x0_dev,
                           const double *u_init_dev, const double *zs_dev, void *workspace_dev, size_t workspace_bytes,
                           const isls_solve_out *out, int32_t ls_ctas, int32_t ff_depth, double *ms_host, void *stream);
/* dependent-free DFMA throughput of the whole GPU in TFLOP/s (FMA = 2 flop), measured with CUDA events */
int isls_measure_fp64_tflops(double *tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ISLS_B200_H */
