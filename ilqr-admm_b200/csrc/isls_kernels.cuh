// isls_kernels.cuh - model-templated kernels and launch sequences of libisls_b200.so.  Included by one translation
// unit per dynamics model (isls_model_*.cu, each instantiating ModelImpl<M> behind an isls_model_ops table) and by
// isls_b200.cu (C-ABI, plans, workspace carving, model-independent kernels), so the models compile in parallel.
//
// Mapping (B200-first, not a port of the numpy code):
//   * one trajectory per thread for the sequential recursions (Riccati K-pass, feed-forward pass, linear
//     rollout, winner rollout + ADMM projection/dual update); the time loop stays in-kernel.
//   * the line search runs the L candidates of 32 problems in one CTA: lane = problem, each thread carries
//     CPT candidates as independent FP64 dependency chains (ILP), argmin through shared memory.
//   * all per-problem trajectories live in HBM in a tile-blocked SoA layout [tile = b/32][t][component][b%32]:
//     a warp owns one tile, every load/store is one fully coalesced 256-byte line, a tile's data is one
//     contiguous stream (TLB / prefetch friendly), FP64 throughout.
//   * problems are independent: per-problem `done` flags implement the reference's stop rules; no collectives.
//
// The algorithm follows SURVEY.md 8(c'): Riccati form of the reference's dense batch least-squares inner solve
// (isls/isls.py:436-465), identical minimiser.  Reference lines are cited at each device function.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/isls_b200.h"
#include "smallmat.cuh"

#define TILE 32
#define MAX_L 50

#include "common.cuh"
#include "soc.cuh"

static inline int fail(int code, const std::string &msg) { return isls_fail(code, msg); }

// ---------------------------------------------------------------------------------------- per-kernel event timing
// Optional (bench.py roofline): when enabled on the calling thread, every kernel launch of a solve is bracketed by
// a CUDA event pair on the launching stream; isls_profile_collect() synchronises and sums per kernel class.
struct ProfRec { int kc; cudaEvent_t e0, e1; };
extern thread_local bool g_prof_on;                 // defined in isls_b200.cu
extern thread_local std::vector<ProfRec> g_prof;
struct ProfScope {
  cudaStream_t s; int idx;
  ProfScope(int kc, cudaStream_t s_) : s(s_), idx(-1) {
    if (!g_prof_on) return;
    ProfRec r; r.kc = kc;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
    g_prof.push_back(r); idx = (int)g_prof.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(g_prof[idx].e1, s); }
};
#define LAUNCH(kc, stream, ...) do { ProfScope ps__(kc, stream); __VA_ARGS__; } while (0)

// Batches below this many tiles (less than one warp per SM scheduler and wave) take the latency-oriented kernel forms:
// Jacobian cache, deep operand rings, plan constants one step ahead.  ISLS_SMALL_TILES overrides (tuning).
static inline int isls_small_tiles() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("ISLS_SMALL_TILES"); v = e ? atoi(e) : 1536; }
  return v;
}

// ------------------------------------------------------------------------------------------------ device context
struct Dev {
  int N, n_via, L, proj_x, proj_u, T;
  int tile0, tile1;      // tile range of this launch
  int tstep;             // tile stride of this launch (k_ff, k_ff_tma, k_linesearch: 2 = every other tile, the half
                         // batches of the overlapped schedule; all other kernels use stride 1)
  long long B;
  double dt, u_std;
  double Rw[3];          // R = u_std * diag(Rw); Rw = 1 unless the plan carries Rdiag (exact for the scalar-R plans)
  int cost_kind;         // ISLS_COST_*
  // plan constants (device)
  const double *qd;      // [N][n]  Qdiag[seq[t]] (pseudo-Huber: weights of term a)
  const double *hp, *qd2, *hp2;   // [N][n] pseudo-Huber smoothness of term a, weights / smoothness of term b
  const int *seq;        // [N]
  const int *qnz;        // [N]  1 if Qdiag[seq[t]] has a non-zero
  const double *rho_x, *lo_x, *hi_x;   // [N][n]
  const double *rho_u, *lo_u, *hi_u;   // [N][m]
  const double *alphas;  // [L]
  // obstacle sets of the state projection (n_obst = 0: box): centres [K][2], W / W^-1 [K][4], lower [K]
  int n_obst, obst_max_iter, obst_kind, obst_dyk_max_iter;
  double obst_upper, obst_rho, obst_threshold, obst_dyk_tol;
  const double *ob_c, *ob_W, *ob_Wi, *ob_lo;
  double *obw;           // workspace [2][T][N][n][32]: winner x, pre-projection point
  // robust iSLS-ADMM: C = dim + 1 columns [d_u | Phi_u(:, :dim)]; Zm, Lm (ADMM z, lambda, delta coordinates) and Xu
  // (primal iterate) are [T][N][m * C][32] with component index j * C + c
  int isls_C, ls_cost_only;
  double *Zm, *Lm, *Xu;
  // state side of isls_admm (project_x, isls.py:631-638): the same three arrays for [d_x | Phi_x(:, :dim)],
  // [T][N][n * C][32] with component index i * C + c; NULL without a state projection
  double *Zx, *Lx, *Xx;
  double stall_tol, osc_tol;
  // workspace, tile-blocked [T][N][dim][32]
  double *xh, *uh, *du, *zx, *lx, *zu, *lu, *rgx, *rgu, *Kg, *Qux, *Quu, *Qui, *kk, *zs;   // Quu, Qui: packed lower
  double *Jc;            // [T][N][NJA][32] Jacobian scalars of the current linearisation (small batches only; else NULL)
  double *lsc;           // [T][L][32] candidate costs of the last line search
  // per-problem scalars [T*32]
  double *cost, *prev_cost, *prim, *dual, *cost_adm, *best_cost;
  double *cq;            // [6][T*32] control-cost polynomials of the current line search: rows 0-2 cost + ADMM
                         // penalty c0 + a c1 + a^2 c2, rows 3-5 the R-only part (cost of the winner without penalty)
  int *best, *odone, *adone, *nlog, *status, *oit, *ait;
  int *orig;             // slot -> original problem index (-1: empty); NULL = identity (no compaction)
  int *newpos, *nact;    // compaction scratch: new slot of every old slot, number of active problems
  // options
  int max_outer, max_admm, fixed_budget, last_stage_dp;
  double tol, outer_tol, relax;
  isls_solve_out out;
};

#define EL(p, dim, t, c) (p)[((size_t)(t) * (dim) + (c)) * TILE]

template <class M>
struct TileCtx {
  int tile, lane;
  long long b, ob;   // slot index (workspace), original problem index (natural-layout inputs / outputs; clamped)
  bool valid;        // the slot holds a real problem
  __device__ __forceinline__ TileCtx(const Dev &d, int tile_, int lane_) : tile(tile_), lane(lane_) {
    b = (long long)tile * TILE + lane;
    if (d.orig) {                       // compacting solves: slots are re-packed between outer iterations
      const int o = d.orig[b];
      valid = o >= 0;
      ob = valid ? o : 0;
    } else {
      valid = b < d.B;
      ob = valid ? b : d.B - 1;
    }
  }
  __device__ __forceinline__ double *at(double *base, const Dev &d, int dim) const {
    return base + (size_t)tile * d.N * dim * TILE + lane;
  }
};

template <class M>
__device__ __forceinline__ void retire(const Dev &d, const TileCtx<M> &c, bool finished);

// State cost of one step.  Quadratic via-point cost sum_i Qd[t][i] (x_i - z_i)^2 (sls_base.py:25-44) or the
// pseudo-Huber family of the Tutorial (notebooks/Tutorial.ipynb cell 14): sum_i w (sqrt(e^2 + p^2) - p), e = x_i - z_i,
// with two (w, p) terms per component (running cost lx and final cost lf both act on x, y at the last step).
__device__ __forceinline__ double huber_val(double e, double w, double p) { return w * (sqrt(fma(e, e, p * p)) - p); }
template <class M>
__device__ __forceinline__ double state_cost(const Dev &d, const double *zs, int t, const double (&x)[M::n],
                                             int qnz_t = -1) {     // qnz_t >= 0: d.qnz[t] already fetched by the caller
  double c = 0.0;
  if (qnz_t >= 0 ? qnz_t : d.qnz[t]) {
    const int s = d.seq[t];
    if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
      for (int i = 0; i < M::n; i++) {
        const double dx = x[i] - EL(zs, M::n, s, i);
        c += (dx * dx) * d.qd[t * M::n + i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < M::n; i++) {
        const double e = x[i] - EL(zs, M::n, s, i);
        const double wa = d.qd[t * M::n + i], wb = d.qd2[t * M::n + i];
        if (wa != 0.0) c += huber_val(e, wa, d.hp[t * M::n + i]);
        if (wb != 0.0) c += huber_val(e, wb, d.hp2[t * M::n + i]);
      }
    }
  }
  return c;
}
// gradient g and Hessian diagonal h of the state cost at x (the reference's cts[:, :n] and diag(Cts[:, :n, :n]),
// isls/isls.py:263-279): quadratic 2Q(x - z), 2Q; pseudo-Huber w e / s, w p^2 / s^3 with s = sqrt(e^2 + p^2).
template <class M>
__device__ __forceinline__ void state_grad_hess(const Dev &d, const double *zs, int t, const double (&x)[M::n],
                                                double (&g)[M::n], double (&h)[M::n]) {
#pragma unroll
  for (int i = 0; i < M::n; i++) g[i] = h[i] = 0.0;
  if (!d.qnz[t]) return;
  const int s = d.seq[t];
  if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
    for (int i = 0; i < M::n; i++) {
      const double q = d.qd[t * M::n + i];
      g[i] = 2.0 * q * (x[i] - EL(zs, M::n, s, i));
      h[i] = 2.0 * q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < M::n; i++) {
      const double e = x[i] - EL(zs, M::n, s, i);
      const double wa = d.qd[t * M::n + i], wb = d.qd2[t * M::n + i];
      if (wa != 0.0) {
        const double pp = d.hp[t * M::n + i], sq = sqrt(fma(e, e, pp * pp));
        g[i] += wa * e / sq;
        h[i] += wa * (pp * pp) / (sq * sq * sq);
      }
      if (wb != 0.0) {
        const double pp = d.hp2[t * M::n + i], sq = sqrt(fma(e, e, pp * pp));
        g[i] += wb * e / sq;
        h[i] += wb * (pp * pp) / (sq * sq * sq);
      }
    }
  }
}
// control cost u'Ru / u_std = sum_j Rw_j u_j^2 (Rw = 1: the multiplication is exact)
template <class M>
__device__ __forceinline__ double ctrl_sq(const Dev &d, const double (&u)[M::m]) {
  double c = 0.0;
#pragma unroll
  for (int j = 0; j < M::m; j++) c += (d.Rw[j] * u[j]) * u[j];
  return c;
}

// ------------------------------------------------------------------------------------------------------ kernels
// Initial rollout of the user's control guess from x0 (what the reference user does through rollout_batch +
// nominal_values, isls/isls.py:135-154, isls/isls_base.py:80-85) + workspace initialisation.
template <class M>
__global__ void k_init(Dev d, const double *x0, const double *u_init, const double *zs_in) {
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  int *orig = d.orig;
  d.orig = nullptr;                       // slots == problems at initialisation
  TileCtx<M> c(d, tile, threadIdx.x);
  if (orig) orig[c.b] = c.valid ? (int)c.b : -1;
  double *xh = c.at(d.xh, d, M::n), *uh = c.at(d.uh, d, M::m);
  double *zs = d.zs + (size_t)tile * d.n_via * M::n * TILE + c.lane;
  for (int k = 0; k < d.n_via; k++)
#pragma unroll
    for (int i = 0; i < M::n; i++) EL(zs, M::n, k, i) = zs_in[(c.ob * d.n_via + k) * M::n + i];
  double x[M::n], u[M::m], xn[M::n];
#pragma unroll
  for (int i = 0; i < M::n; i++) x[i] = x0[c.ob * M::n + i];
  double cs = 0.0, cc = 0.0;
  double *zx = c.at(d.zx, d, M::n), *zu = c.at(d.zu, d, M::m);
  double *lx = c.at(d.lx, d, M::n), *lu = c.at(d.lu, d, M::m);
  for (int t = 0; t < d.N; t++) {
#pragma unroll
    for (int j = 0; j < M::m; j++) {
      u[j] = u_init[(c.ob * d.N + t) * M::m + j];
      EL(uh, M::m, t, j) = u[j];
      EL(zu, M::m, t, j) = 0.0;
      EL(lu, M::m, t, j) = 0.0;
    }
    cc += ctrl_sq<M>(d, u);
#pragma unroll
    for (int i = 0; i < M::n; i++) {
      EL(xh, M::n, t, i) = x[i];
      EL(zx, M::n, t, i) = 0.0;
      EL(lx, M::n, t, i) = 0.0;
    }
    cs += state_cost<M>(d, zs, t, x);
    M::step(x, u, xn, d.dt);
#pragma unroll
    for (int i = 0; i < M::n; i++) x[i] = xn[i];
    if (d.isls_C > 0) {                                  // z_u_init = 0 (isls.py:537)
      double *Zm = c.at(d.Zm, d, M::m * d.isls_C);
      for (int q = 0; q < M::m * d.isls_C; q++) EL(Zm, M::m * d.isls_C, t, q) = 0.0;
      if (d.Zx) {                                        // z_x_init = 0 (isls.py:536)
        double *Zx = c.at(d.Zx, d, M::n * d.isls_C);
        for (int q = 0; q < M::n * d.isls_C; q++) EL(Zx, M::n * d.isls_C, t, q) = 0.0;
      }
    }
  }
  const double cost = cs + d.u_std * cc;
  d.cost[c.b] = cost;
  d.prev_cost[c.b] = cost;
  d.nlog[c.b] = 1;
  d.status[c.b] = 0;
  d.odone[c.b] = 0;
  d.adone[c.b] = 0;
  d.oit[c.b] = 0;
  d.ait[c.b] = 0;
  d.best[c.b] = 0;
  if (c.valid) {
    double *cl = d.out.cost_log + c.ob * (d.max_outer + 1);
    cl[0] = cost;
    for (int i = 1; i <= d.max_outer; i++) cl[i] = nan("");
    if (d.out.admm_iters) for (int i = 0; i < d.max_outer; i++) d.out.admm_iters[c.ob * d.max_outer + i] = 0;
    if (d.out.admm_exit) for (int i = 0; i < d.max_outer; i++) d.out.admm_exit[c.ob * d.max_outer + i] = 0;
    if (d.out.res_log)
      for (int i = 0; i < d.max_outer * d.max_admm * 2; i++)
        d.out.res_log[(size_t)c.ob * d.max_outer * d.max_admm * 2 + i] = nan("");
    if (d.out.alpha_idx)
      for (int i = 0; i < d.max_outer * d.max_admm; i++)
        d.out.alpha_idx[(size_t)c.ob * d.max_outer * d.max_admm + i] = -1;
    if (d.out.inner_iters)
      for (int i = 0; i < d.max_outer * d.max_admm; i++)
        d.out.inner_iters[(size_t)c.ob * d.max_outer * d.max_admm + i] = 0;
  }
}

// One Riccati step shared by the K-pass (ADMM path) and the full backward pass (plain iLQR):
//   Qxx = Cxx + A'VA, Qux = B'VA, Quu = Cuu + B'VB           isls/isls.py:288-290 (Cux = 0, diagonal Cxx/Cuu)
//   K = -Quu^-1 Qux                                           isls/isls.py:296-297 (sls.py:149-150 form)
//   V = Qxx + K'QuuK + Qux'K + K'Qux                          isls/isls.py:300
// JOSEPH: the same V in the closed-loop ("Joseph") form  V = Cxx + K'Cuu K + (A + BK)'V(A + BK)  - a sum of
// positive semi-definite terms instead of the four-term expression above, whose cancellation is unstable in FP64 when
// the control is cheap next to the accumulated state weights (Cuu << B'VB: with R = 1e-4 and a state penalty Qr = 10
// on the joint velocities of the arm the four-term recursion loses all but 1-2 digits of the first controls against
// the exact minimiser, the Joseph form keeps 1e-11; oracle/restated.py::backward_pass).  Used where the reference
// solves the dense normal equations with such weights (isls_admm); the four-term form mirrors the reference's own DP
// recursion and stays on every other path.
template <class M, bool JOSEPH = false>
__device__ __forceinline__ bool riccati_step(const double (&A)[M::n][M::n], const double (&Bm)[M::n][M::m],
                                             const double (&dxx)[M::n], const double (&duu)[M::m],
                                             double (&V)[M::n][M::n], double (&K)[M::m][M::n],
                                             double (&Qux)[M::m][M::n], double (&Quu)[M::m][M::m],
                                             double (&Qui)[M::m][M::m]) {
  constexpr int n = M::n, m = M::m;
  double VA[n][n], Qxx[n][n];
  mat_V_A<M>(V, A, VA);
  mat_At_X<M, n>(A, VA, Qxx);
#pragma unroll
  for (int i = 0; i < n; i++) Qxx[i][i] += dxx[i];
  mat_Bt_X<M, n>(Bm, VA, Qux);
  double VB[n][m];
  mat_V_B<M>(V, Bm, VB);
  mat_Bt_X<M, m>(Bm, VB, Quu);
#pragma unroll
  for (int i = 0; i < m; i++) Quu[i][i] += duu[i];
  const bool ok = spd_inverse<m>(Quu, Qui);
#pragma unroll
  for (int a = 0; a < m; a++)
#pragma unroll
    for (int j = 0; j < n; j++) {
      double acc = 0.0;
#pragma unroll
      for (int b2 = 0; b2 < m; b2++) acc = fma(Qui[a][b2], Qux[b2][j], acc);
      K[a][j] = -acc;
    }
  if (JOSEPH) {
    // W = V (A + BK) = VA + VB K;  V' = diag(dxx) + K' diag(duu) K + A'W + K'(B'W), symmetrised
    double W[n][n], AtW[n][n], BtW[m][n];
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
      for (int j = 0; j < n; j++) {
        double acc = VA[i][j];
#pragma unroll
        for (int a = 0; a < m; a++) acc = fma(VB[i][a], K[a][j], acc);
        W[i][j] = acc;
      }
    mat_At_X<M, n>(A, W, AtW);
    mat_Bt_X<M, n>(Bm, W, BtW);
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
      for (int j = 0; j < n; j++) {
        double acc = AtW[i][j];
#pragma unroll
        for (int a = 0; a < m; a++) acc = fma(K[a][i], fma(duu[a], K[a][j], BtW[a][j]), acc);
        Qxx[i][j] = acc;
      }
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
      for (int j = 0; j < n; j++) V[i][j] = 0.5 * (Qxx[i][j] + Qxx[j][i]) + (i == j ? dxx[i] : 0.0);
    return ok;
  }
  double QK[m][n];
#pragma unroll
  for (int a = 0; a < m; a++)
#pragma unroll
    for (int j = 0; j < n; j++) {
      double acc = 0.0;
#pragma unroll
      for (int b2 = 0; b2 < m; b2++) acc = fma(Quu[a][b2], K[b2][j], acc);
      QK[a][j] = acc;
    }
#pragma unroll
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int j = 0; j < n; j++) {
      double t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
      for (int a = 0; a < m; a++) {
        t1 = fma(K[a][i], QK[a][j], t1);
        t2 = fma(Qux[a][i], K[a][j], t2);
        t3 = fma(K[a][i], Qux[a][j], t3);
      }
      V[i][j] = ((Qxx[i][j] + t1) + t2) + t3;
    }
  return ok;
}

// feed-forward step (sls.py:196-199): qx = cx + A'v, qu = cu + B'v, k = -Quu^-1 qu,
//   v = qx + Qux'k + K'qu + K'(Quu k).
// The reference's four-term form is v = qx + Qux'k + K'w with the residual w = qu + Quu k: the K'w term cancels the
// first-order error of the computed inverse (with K = -Quu^-1 Qux, K'w = -Qux' Quu^-1 w).  The same correction is
// applied here to k itself - one refinement step k <- k - Quu^-1 w - after which v = qx + Qux'k.  It is the
// identical first-order-exact value, needs Qux, Quu, Quu^-1 (no K) in the backward sweep, and hands the rollout a
// refined k.  (The plain collapsed form qx + K'qu is cheaper but first-order sensitive to cond(Quu) - measurably
// worse on the arm.)
template <class M>
__device__ __forceinline__ void ff_step(const double (&A)[M::n][M::n], const double (&Bm)[M::n][M::m],
                                        const double (&cx)[M::n], const double (&cu)[M::m],
                                        const double (&Qux)[M::m][M::n], const double (&Quu)[M::m][M::m],
                                        const double (&Qui)[M::m][M::m], double (&v)[M::n], double (&kt)[M::m]) {
  constexpr int n = M::n, m = M::m;
  double qx[n], qu[m], k0[m], w[m];
  mat_At_v<M>(A, v, qx);
  mat_Bt_v<M>(Bm, v, qu);
#pragma unroll
  for (int i = 0; i < n; i++) qx[i] += cx[i];
#pragma unroll
  for (int j = 0; j < m; j++) qu[j] += cu[j];
#pragma unroll
  for (int a = 0; a < m; a++) {
    double acc = 0.0;
#pragma unroll
    for (int b2 = 0; b2 < m; b2++) acc = fma(Qui[a][b2], qu[b2], acc);
    k0[a] = -acc;
  }
#pragma unroll
  for (int a = 0; a < m; a++) {
    double acc = qu[a];
#pragma unroll
    for (int b2 = 0; b2 < m; b2++) acc = fma(Quu[a][b2], k0[b2], acc);
    w[a] = acc;
  }
#pragma unroll
  for (int a = 0; a < m; a++) {
    double acc = 0.0;
#pragma unroll
    for (int b2 = 0; b2 < m; b2++) acc = fma(Qui[a][b2], w[b2], acc);
    kt[a] = k0[a] - acc;
  }
#pragma unroll
  for (int i = 0; i < n; i++) {
    double acc = qx[i];
#pragma unroll
    for (int a = 0; a < m; a++) acc = fma(Qux[a][i], kt[a], acc);
    v[i] = acc;
  }
}

// packed lower triangle of a symmetric m x m matrix: index of (a, b), b <= a
__host__ __device__ constexpr int tri(int a, int b) { return a * (a + 1) / 2 + b; }
#define NTRI(m) ((m) * ((m) + 1) / 2)

template <class M>
__device__ __forceinline__ void init_AB(double (&A)[M::n][M::n], double (&Bm)[M::n][M::m]) {
#pragma unroll
  for (int i = 0; i < M::n; i++) {
#pragma unroll
    for (int j = 0; j < M::n; j++) A[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < M::m; j++) Bm[i][j] = 0.0;
  }
}

// K-pass: linearise at the nominal trajectory (get_AB, isls/isls.py:424) and run the Riccati recursion with
// Cxx = 2(Q_t + Qr_t), Cuu = 2(R + Rr_t) (regularised form of sls.py:132-137).  Stores K_t, Qux_t and the packed
// Quu_t, Quu_t^-1 (the logs of sls.py:159-162 that the feed-forward passes need); resets the ADMM state of the new outer
// iteration (lambda = 0, isls/isls.py:414-415; z warm start isls/isls.py:489-490).
template <class M, bool SMALL = false, bool JOSEPH = false>
__global__ void k_kpass(Dev d) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m);
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b]) return;
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  double *Kg = c.at(d.Kg, d, m * n), *Qx = c.at(d.Qux, d, m * n);
  double *Qu = c.at(d.Quu, d, nt), *Qi = c.at(d.Qui, d, nt);
  double A[n][n], Bm[n][m], V[n][n];
  init_AB<M>(A, Bm);
#pragma unroll
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int j = 0; j < n; j++)
      V[i][j] = 0.0;
  {
    double xl[n], gl[n], hl[n];
#pragma unroll
    for (int i = 0; i < n; i++) xl[i] = (d.cost_kind != ISLS_COST_QUADRATIC) ? EL(xh, n, d.N - 1, i) : 0.0;
    state_grad_hess<M>(d, zs, d.N - 1, xl, gl, hl);
#pragma unroll
    for (int i = 0; i < n; i++) V[i][i] = hl[i] + 2.0 * d.rho_x[(d.N - 1) * n + i];              // isls.py:257
  }
  bool ok = true;
  double *Jc = d.Jc ? c.at(d.Jc, d, M::NJA) : nullptr;
  // operands of the next step, loaded while this step's Riccati update runs - including the plan constants (q_t, rho_t:
  // as plain loads inside the step they were an exposed L2 round trip per step, a quarter of this kernel's stall
  // samples at 8,192 problems; profiles/r2_small_batch.md)
  // (SMALL only: at 65,536 problems the kernel is HBM-bound and the 34 extra registers cost residency - 0.31 -> 0.41 ms)
  double xn_[n], un_[m], qdn_[SMALL ? n : 1], rxn_[SMALL ? n : 1], run_[SMALL ? m : 1];
  auto prefetch = [&](int t) {
#pragma unroll
    for (int i = 0; i < n; i++) {
      xn_[i] = EL(xh, n, t, i);
      if (SMALL) { qdn_[i] = __ldg(d.qd + t * n + i); rxn_[i] = __ldg(d.rho_x + t * n + i); }
    }
#pragma unroll
    for (int j = 0; j < m; j++) {
      un_[j] = EL(uh, m, t, j);
      if (SMALL) run_[j] = __ldg(d.rho_u + t * m + j);
    }
  };
  prefetch(d.N - 2);
  for (int t = d.N - 2; t >= 0; t--) {
    double x[n], u[m], J[M::NJA], qd_t[n], rx_t[n], ru_t[m];
#pragma unroll
    for (int i = 0; i < n; i++) {
      x[i] = xn_[i];
      qd_t[i] = SMALL ? qdn_[i] : d.qd[t * n + i];
      rx_t[i] = SMALL ? rxn_[i] : d.rho_x[t * n + i];
    }
#pragma unroll
    for (int j = 0; j < m; j++) { u[j] = un_[j]; ru_t[j] = SMALL ? run_[j] : d.rho_u[t * m + j]; }
    if (t > 0) prefetch(t - 1);
    M::jac(x, u, J, d.dt);
    M::expand(J, A, Bm, d.dt);
    if (SMALL && Jc) {
#pragma unroll
      for (int q = 0; q < M::NJA; q++) EL(Jc, M::NJA, t, q) = J[q];
    }
    double dxx[n], duu[m], K[m][n], Qux[m][n], Quu[m][m], Qui[m][m];
    if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
      for (int i = 0; i < n; i++) dxx[i] = 2.0 * (qd_t[i] + rx_t[i]);
    } else {
      double gt[n], ht[n];
      state_grad_hess<M>(d, zs, t, x, gt, ht);
#pragma unroll
      for (int i = 0; i < n; i++) dxx[i] = ht[i] + 2.0 * rx_t[i];
    }
#pragma unroll
    for (int j = 0; j < m; j++) duu[j] = 2.0 * (d.u_std * d.Rw[j] + ru_t[j]);
    ok &= riccati_step<M, JOSEPH>(A, Bm, dxx, duu, V, K, Qux, Quu, Qui);
#pragma unroll
    for (int a = 0; a < m; a++) {
#pragma unroll
      for (int j = 0; j < n; j++) { EL(Kg, m * n, t, a * n + j) = K[a][j]; EL(Qx, m * n, t, a * n + j) = Qux[a][j]; }
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) {
        EL(Qu, nt, t, tri(a, b2)) = 0.5 * (Quu[a][b2] + Quu[b2][a]);
        EL(Qi, nt, t, tri(a, b2)) = Qui[a][b2];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < m * n; q++) EL(Kg, m * n, d.N - 1, q) = 0.0;
  // reset ADMM state for this outer iteration
  if (d.proj_x) {
    double *zx = c.at(d.zx, d, n), *lx = c.at(d.lx, d, n), *rg = c.at(d.rgx, d, n);
    for (int t0 = 0; t0 < d.N; t0 += 4) {            // 4 steps of loads in flight before the first store
      double zv[4][n];
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int i = 0; i < n; i++) zv[q][i] = (t0 + q < d.N) ? EL(zx, n, t0 + q, i) : 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (t0 + q < d.N) {
#pragma unroll
          for (int i = 0; i < n; i++) { EL(lx, n, t0 + q, i) = 0.0; EL(rg, n, t0 + q, i) = zv[q][i]; }
        }
    }
  }
  if (d.proj_u) {
    double *zu = c.at(d.zu, d, m), *lu = c.at(d.lu, d, m), *rg = c.at(d.rgu, d, m);
    for (int t0 = 0; t0 < d.N; t0 += 8) {
      double zv[8][m];
#pragma unroll
      for (int q = 0; q < 8; q++)
#pragma unroll
        for (int j = 0; j < m; j++) zv[q][j] = (t0 + q < d.N) ? EL(zu, m, t0 + q, j) : 0.0;
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (t0 + q < d.N) {
#pragma unroll
          for (int j = 0; j < m; j++) { EL(lu, m, t0 + q, j) = 0.0; EL(rg, m, t0 + q, j) = zv[q][j]; }
        }
    }
  }
  if (!ok) d.status[c.b] |= ISLS_ST_NON_PD;
  d.prev_cost[c.b] = d.cost[c.b];
  d.adone[c.b] = 0;
  d.ait[c.b] = 0;
  d.prim[c.b] = 1e6;     // admm.py:24-25
  d.dual[c.b] = 1e6;
}

// Cost-gradient terms of the ff-pass with the rounding order spelled out (no compiler-chosen FMA contraction): the
// plain, cp.async-staged and TMA-staged variants must produce the same bits (tests/test_gpu_kernel_variants.py; the
// a*b + c*d form left the choice of the fused product to the compiler and the arm differed by one ulp between variants).
__device__ __forceinline__ double ff_cx_quad(double qd, double x, double z) {            // 2 Q (x^ - z_via)
  return __dmul_rn(__dmul_rn(2.0, qd), __dsub_rn(x, z));
}
__device__ __forceinline__ double ff_pen(double c, double rho, double v, double reg) {   // c + 2 rho (v - reg)
  return __fma_rn(__dmul_rn(2.0, rho), __dsub_rn(v, reg), c);
}

// ff-pass + linear rollout = the argmin of the regularised LQ problem, delta_u* (isls/isls.py:457-465 in
// Riccati form, SURVEY 8c' step 2): backward feed-forward recursion (sls.py:168-202) with
//   cx = 2Q(x^ - z_via) + 2Qr(x^ - reg_x), cu = 2R u^ + 2Rr(u^ - reg_u),
// batch-form last control, then du_t = K dx + k, dx+ = A dx + B du.
// HBM-bound kernel: the Jacobian scalars are recomputed from (x^_t, u^_t) (loaded anyway for cx, cu) instead of
// being read back; the backward sweep streams Qux_t and the packed Quu_t, Quu_t^-1, the forward sweep K_t.
template <class M>
__device__ __forceinline__ void ff_body(const Dev &d, const TileCtx<M> &c) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m);
  const int tile = c.tile;
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *Kg = c.at(d.Kg, d, m * n), *Qx = c.at(d.Qux, d, m * n);
  const double *Qu = c.at(d.Quu, d, nt), *Qi = c.at(d.Qui, d, nt);
  const double *rgx = c.at(d.rgx, d, n), *rgu = c.at(d.rgu, d, m);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  double *kk = c.at(d.kk, d, m), *du = c.at(d.du, d, m);
  double A[n][n], Bm[n][m];
  init_AB<M>(A, Bm);
  double v[n];
  auto costgrad = [&](int t, const double (&x)[n], const double (&u)[m], double (&cx)[n], double (&cu)[m]) {
    if (d.cost_kind == ISLS_COST_QUADRATIC) {
      const int s = d.seq[t];
#pragma unroll
      for (int i = 0; i < n; i++) cx[i] = d.qnz[t] ? ff_cx_quad(d.qd[t * n + i], x[i], EL(zs, n, s, i)) : 0.0;
    } else {
      double ht[n];
      state_grad_hess<M>(d, zs, t, x, cx, ht);
    }
#pragma unroll
    for (int i = 0; i < n; i++) {
      if (d.proj_x) cx[i] = ff_pen(cx[i], d.rho_x[t * n + i], x[i], EL(rgx, n, t, i));
    }
#pragma unroll
    for (int j = 0; j < m; j++) {
      double g = __dmul_rn(2.0 * (d.u_std * d.Rw[j]), u[j]);
      if (d.proj_u) g = ff_pen(g, d.rho_u[t * m + j], u[j], EL(rgu, m, t, j));
      cu[j] = g;
    }
  };
  {
    double x[n], u[m], cx[n], cu[m];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, d.N - 1, i);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = EL(uh, m, d.N - 1, j);
    costgrad(d.N - 1, x, u, cx, cu);
#pragma unroll
    for (int i = 0; i < n; i++) v[i] = cx[i];
#pragma unroll
    for (int j = 0; j < m; j++) {
      // batch-form last control: du_{N-1} = -Cuu^-1 cu (isls.py:441-465, Su's last block column is zero)
      const double cuu = __dmul_rn(2.0, __fma_rn(d.u_std, d.Rw[j], d.rho_u[(d.N - 1) * m + j]));
      EL(kk, m, d.N - 1, j) = d.last_stage_dp ? 0.0 : -cu[j] / cuu;
    }
  }
  for (int t = d.N - 2; t >= 0; t--) {
    double x[n], u[m], J[M::NJA], cx[n], cu[m], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, t, i);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = EL(uh, m, t, j);
#pragma unroll
    for (int a = 0; a < m; a++) {
#pragma unroll
      for (int j = 0; j < n; j++) Qux[a][j] = EL(Qx, m * n, t, a * n + j);
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) {
        Qui[a][b2] = EL(Qi, nt, t, tri(a, b2)); Qui[b2][a] = Qui[a][b2];
        Quu[a][b2] = EL(Qu, nt, t, tri(a, b2)); Quu[b2][a] = Quu[a][b2];
      }
    }
    M::jac(x, u, J, d.dt);
    M::expand(J, A, Bm, d.dt);
    costgrad(t, x, u, cx, cu);
    ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
#pragma unroll
    for (int j = 0; j < m; j++) EL(kk, m, t, j) = kt[j];
  }
  // linear rollout.  The control part of every line-search candidate's cost,
  //   sum_t R u^2 + Rr (u - reg_u)^2  with  u = u^ + alpha du          (isls.py:470, 474-476)
  // is the quadratic c0 + alpha c1 + alpha^2 c2 in alpha; its three coefficients are accumulated here once per
  // problem so the L candidate rollouts neither recompute it nor read reg_u.
  double dx[n];
#pragma unroll
  for (int i = 0; i < n; i++) dx[i] = 0.0;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
  for (int t = 0; t < d.N; t++) {
    // every load of the step is issued before its first store (see admm_body)
    double duv[m], u[m], K[m][n], kv[m], ru[m], x[n];
#pragma unroll
    for (int a = 0; a < m; a++) {
#pragma unroll
      for (int j = 0; j < n; j++) K[a][j] = EL(Kg, m * n, t, a * n + j);
      kv[a] = EL(kk, m, t, a);
      u[a] = EL(uh, m, t, a);
      ru[a] = d.proj_u ? EL(rgu, m, t, a) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, t, i);
#pragma unroll
    for (int a = 0; a < m; a++) {
      double acc = 0.0;
      if (t < d.N - 1) {
#pragma unroll
        for (int j = 0; j < n; j++) acc = fma(K[a][j], dx[j], acc);
      }
      duv[a] = acc + kv[a];
      EL(du, m, t, a) = duv[a];
      r0 = fma(d.Rw[a] * u[a], u[a], r0);       // R-only part (scaled by u_std after the loop)
      r1 = fma(d.Rw[a] * u[a], duv[a], r1);
      r2 = fma(d.Rw[a] * duv[a], duv[a], r2);
      if (d.proj_u) {                           // ADMM penalty part
        const double rho = d.rho_u[t * m + a], e = u[a] - ru[a];
        c0 = fma(rho * e, e, c0);
        c1 = fma(2.0 * rho * e, duv[a], c1);
        c2 = fma(rho * duv[a], duv[a], c2);
      }
    }
    if (t < d.N - 1) {
      double J[M::NJA], dxn[n];
      M::jac(x, u, J, d.dt);
      M::expand(J, A, Bm, d.dt);
      mat_Ax_Bu<M>(A, Bm, dx, duv, dxn);
#pragma unroll
      for (int i = 0; i < n; i++) dx[i] = dxn[i];
    }
  }
  const size_t S = (size_t)d.T * TILE;
  r0 *= d.u_std;                       // R sum u^2, 2R sum u du, R sum du^2  (R = u_std I)
  r1 *= 2.0 * d.u_std;
  r2 *= d.u_std;
  d.cq[c.b] = c0 + r0;
  d.cq[S + c.b] = c1 + r1;
  d.cq[2 * S + c.b] = c2 + r2;
  d.cq[3 * S + c.b] = r0;
  d.cq[4 * S + c.b] = r1;
  d.cq[5 * S + c.b] = r2;
}

template <class M>
__global__ void k_ff(Dev d) {
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b] || d.adone[c.b]) return;
  ff_body<M>(d, c);
}

// ---- cp.async (LDGSTS) staging: each thread streams ITS OWN operands of the next STAGES-1 time steps into a private
// slice of shared memory, so several steps of loads are in flight per warp without holding registers.  Used for the
// ff-pass when the batch is too small to hide the per-step memory round trip with resident warps alone (at 65,536
// problems the plain kernel is bandwidth-bound and staging is neutral; at 4,096 - 16,384 it is latency-bound).
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int Npend>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(Npend) : "memory"); }

template <class M, int STAGES>
__global__ void __launch_bounds__(TILE) k_ff_staged(Dev d) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m);
  constexpr int SB = n + m + m * n + 2 * nt + m + n;      // backward slots: x^, u^, Qux, Quu, Quu^-1, reg_u, reg_x
  constexpr int SF = m * n + m + m + m + n;               // forward slots:  K, k, u^, reg_u, x^
  constexpr int SL = SB > SF ? SB : SF;
  extern __shared__ double smem_ff[];
  const int tile = d.tile0 + blockIdx.x;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b] || d.adone[c.b]) return;
  const int tid = threadIdx.x;
  auto slot = [&](int stage, int k) -> double * { return smem_ff + ((size_t)stage * SL + k) * TILE + tid; };
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *Kg = c.at(d.Kg, d, m * n), *Qx = c.at(d.Qux, d, m * n);
  const double *Qu = c.at(d.Quu, d, nt), *Qi = c.at(d.Qui, d, nt);
  const double *rgx = c.at(d.rgx, d, n), *rgu = c.at(d.rgu, d, m);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  double *kk = c.at(d.kk, d, m), *du = c.at(d.du, d, m);
  const int N = d.N;
  double A[n][n], Bm[n][m];
  init_AB<M>(A, Bm);
  double v[n];
  // Plan constants of a step (rho_x, rho_u, seq, qnz) are fetched ONE STEP AHEAD into registers (ldc, read-only
  // path; re-loaded in place right after their last use in costgrad, so there is no second register set): as plain
  // loads inside the step they sat behind the previous step's stores and missed L1 (the cp.async.ca stream flushes
  // it), a third of this kernel's stall samples at C3 (profiles/r1_c3_small_batch_kernels.md).
  double rw2[m], rw[m];
#pragma unroll
  for (int j = 0; j < m; j++) { rw[j] = d.Rw[j]; rw2[j] = 2.0 * (d.u_std * rw[j]); }
  auto ldc = [&](int t, double (&rhx)[n], double (&rhu)[m], int &qz, int &sq) {
    qz = __ldg(d.qnz + t);
    sq = __ldg(d.seq + t);
#pragma unroll
    for (int i = 0; i < n; i++) rhx[i] = d.proj_x ? __ldg(d.rho_x + t * n + i) : 0.0;
#pragma unroll
    for (int j = 0; j < m; j++) rhu[j] = d.proj_u ? __ldg(d.rho_u + t * m + j) : 0.0;
  };
  auto costgrad = [&](int t, const double (&x)[n], const double (&u)[m], const double (&rx)[n],
                      const double (&ru)[m], double (&cx)[n], double (&cu)[m], const double (&rhx)[n],
                      const double (&rhu)[m], int qz, int sq) {
    if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
      for (int i = 0; i < n; i++) cx[i] = qz ? ff_cx_quad(d.qd[t * n + i], x[i], EL(zs, n, sq, i)) : 0.0;
    } else {
      double ht[n];
      state_grad_hess<M>(d, zs, t, x, cx, ht);
    }
#pragma unroll
    for (int i = 0; i < n; i++) {
      if (d.proj_x) cx[i] = ff_pen(cx[i], rhx[i], x[i], rx[i]);
    }
#pragma unroll
    for (int j = 0; j < m; j++) {
      double g = __dmul_rn(rw2[j], u[j]);
      if (d.proj_u) g = ff_pen(g, rhu[j], u[j], ru[j]);
      cu[j] = g;
    }
  };
  auto issue_b = [&](int t, int stage) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < n; i++) cp_async8(slot(stage, k++), &EL(xh, n, t, i));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slot(stage, k++), &EL(uh, m, t, j));
#pragma unroll
    for (int q = 0; q < m * n; q++) cp_async8(slot(stage, k++), &EL(Qx, m * n, t, q));
#pragma unroll
    for (int q = 0; q < nt; q++) cp_async8(slot(stage, k++), &EL(Qu, nt, t, q));
#pragma unroll
    for (int q = 0; q < nt; q++) cp_async8(slot(stage, k++), &EL(Qi, nt, t, q));
    if (d.proj_u) {
#pragma unroll
      for (int j = 0; j < m; j++) cp_async8(slot(stage, k + j), &EL(rgu, m, t, j));
    }
    k += m;
    if (d.proj_x) {
#pragma unroll
      for (int i = 0; i < n; i++) cp_async8(slot(stage, k + i), &EL(rgx, n, t, i));
    }
  };
  {   // terminal step N-1 (plain loads)
    double x[n], u[m], rx[n], ru[m], cx[n], cu[m];
#pragma unroll
    for (int i = 0; i < n; i++) { x[i] = EL(xh, n, N - 1, i); rx[i] = d.proj_x ? EL(rgx, n, N - 1, i) : 0.0; }
#pragma unroll
    for (int j = 0; j < m; j++) { u[j] = EL(uh, m, N - 1, j); ru[j] = d.proj_u ? EL(rgu, m, N - 1, j) : 0.0; }
    double rhx[n], rhu[m];
    int qz, sq;
    ldc(N - 1, rhx, rhu, qz, sq);
    costgrad(N - 1, x, u, rx, ru, cx, cu, rhx, rhu, qz, sq);
#pragma unroll
    for (int i = 0; i < n; i++) v[i] = cx[i];
#pragma unroll
    for (int j = 0; j < m; j++) {
      const double cuu = __dmul_rn(2.0, __fma_rn(d.u_std, rw[j], d.rho_u[(N - 1) * m + j]));
      EL(kk, m, N - 1, j) = d.last_stage_dp ? 0.0 : -cu[j] / cuu;
    }
  }
  // ---- backward sweep, STAGES deep
  int t_issue = N - 2;
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (t_issue >= 0) issue_b(t_issue, (N - 2 - t_issue) % STAGES);
    cp_async_commit();
    t_issue--;
  }
  double rhx_c[n], rhu_c[m];
  int qz_c = 0, sq_c = 0;
  if (N >= 2) ldc(N - 2, rhx_c, rhu_c, qz_c, sq_c);
  for (int t = N - 2; t >= 0; t--) {
    if (t_issue >= 0) issue_b(t_issue, (N - 2 - t_issue) % STAGES);
    cp_async_commit();
    t_issue--;
    cp_async_wait<STAGES - 1>();
    const int st = (N - 2 - t) % STAGES;
    double x[n], u[m], rx[n], ru[m], J[M::NJA], cx[n], cu[m], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
    int k = 0;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = *slot(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = *slot(st, k++);
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int j = 0; j < n; j++) Qux[a][j] = *slot(st, k++);
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) { Quu[a][b2] = *slot(st, k + tri(a, b2)); Quu[b2][a] = Quu[a][b2]; }
    k += nt;
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) { Qui[a][b2] = *slot(st, k + tri(a, b2)); Qui[b2][a] = Qui[a][b2]; }
    k += nt;
#pragma unroll
    for (int j = 0; j < m; j++) ru[j] = d.proj_u ? *slot(st, k + j) : 0.0;
    k += m;
#pragma unroll
    for (int i = 0; i < n; i++) rx[i] = d.proj_x ? *slot(st, k + i) : 0.0;
    M::jac(x, u, J, d.dt);
    M::expand(J, A, Bm, d.dt);
    costgrad(t, x, u, rx, ru, cx, cu, rhx_c, rhu_c, qz_c, sq_c);
    if (t > 0) ldc(t - 1, rhx_c, rhu_c, qz_c, sq_c);
    ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
#pragma unroll
    for (int j = 0; j < m; j++) EL(kk, m, t, j) = kt[j];
  }
  cp_async_wait<0>();
  // ---- forward sweep (linear rollout + control-cost polynomials), STAGES deep
  // (the forward step needs fewer slots and is shorter than the backward one: the same shared memory holds SFW >=
  // STAGES forward stages - with only STAGES the operands were not there yet when the step started, 10 % of the
  // kernel's stall samples on the cp.async wait)
  constexpr int SFW = (STAGES * SL) / SF < 8 ? (STAGES * SL) / SF : 8;
  auto slotf = [&](int stage, int k) -> double * { return smem_ff + ((size_t)stage * SF + k) * TILE + tid; };
  auto issue_f = [&](int t, int stage) {
    int k = 0;
#pragma unroll
    for (int q = 0; q < m * n; q++) cp_async8(slotf(stage, k++), &EL(Kg, m * n, t, q));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slotf(stage, k++), &EL(kk, m, t, j));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slotf(stage, k++), &EL(uh, m, t, j));
    if (d.proj_u) {
#pragma unroll
      for (int j = 0; j < m; j++) cp_async8(slotf(stage, k + j), &EL(rgu, m, t, j));
    }
    k += m;
#pragma unroll
    for (int i = 0; i < n; i++) cp_async8(slotf(stage, k++), &EL(xh, n, t, i));
  };
  double dx[n];
#pragma unroll
  for (int i = 0; i < n; i++) dx[i] = 0.0;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
  t_issue = 0;
#pragma unroll
  for (int s = 0; s < SFW - 1; s++) {
    if (t_issue < N) issue_f(t_issue, t_issue % SFW);
    cp_async_commit();
    t_issue++;
  }
  double rho_c[m];
#pragma unroll
  for (int j = 0; j < m; j++) rho_c[j] = d.proj_u ? __ldg(d.rho_u + j) : 0.0;
  for (int t = 0; t < N; t++) {
    if (t_issue < N) issue_f(t_issue, t_issue % SFW);
    cp_async_commit();
    t_issue++;
    cp_async_wait<SFW - 1>();
    const int st = t % SFW;
    double duv[m], u[m], K[m][n], kv[m], ru[m], x[n];
    int k = 0;
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int j = 0; j < n; j++) K[a][j] = *slotf(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) kv[j] = *slotf(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = *slotf(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) ru[j] = d.proj_u ? *slotf(st, k + j) : 0.0;
    k += m;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = *slotf(st, k++);
#pragma unroll
    for (int a = 0; a < m; a++) {
      double acc = 0.0;
      if (t < N - 1) {
#pragma unroll
        for (int j = 0; j < n; j++) acc = fma(K[a][j], dx[j], acc);
      }
      duv[a] = acc + kv[a];
      EL(du, m, t, a) = duv[a];
      r0 = fma(rw[a] * u[a], u[a], r0);
      r1 = fma(rw[a] * u[a], duv[a], r1);
      r2 = fma(rw[a] * duv[a], duv[a], r2);
      if (d.proj_u) {
        const double rho = rho_c[a], e = u[a] - ru[a];
        c0 = fma(rho * e, e, c0);
        c1 = fma(2.0 * rho * e, duv[a], c1);
        c2 = fma(rho * duv[a], duv[a], c2);
      }
    }
#pragma unroll
    for (int j = 0; j < m; j++) rho_c[j] = (d.proj_u && t + 1 < N) ? __ldg(d.rho_u + (t + 1) * m + j) : 0.0;
    if (t < N - 1) {
      double J[M::NJA], dxn[n];
      M::jac(x, u, J, d.dt);
      M::expand(J, A, Bm, d.dt);
      mat_Ax_Bu<M>(A, Bm, dx, duv, dxn);
#pragma unroll
      for (int i = 0; i < n; i++) dx[i] = dxn[i];
    }
  }
  cp_async_wait<0>();
  const size_t S = (size_t)d.T * TILE;
  r0 *= d.u_std;
  r1 *= 2.0 * d.u_std;
  r2 *= d.u_std;
  d.cq[c.b] = c0 + r0;
  d.cq[S + c.b] = c1 + r1;
  d.cq[2 * S + c.b] = c2 + r2;
  d.cq[3 * S + c.b] = r0;
  d.cq[4 * S + c.b] = r1;
  d.cq[5 * S + c.b] = r2;
}

// np.argmin semantics over candidate costs (first minimum; the first NaN wins, isls/isls.py:477)
__device__ __forceinline__ int argmin_np(const double *c, int L, int stride, bool *has_nan) {
  double best = c[0];
  int idx = 0;
  *has_nan = false;
  if (best != best) { *has_nan = true; return 0; }
  for (int l = 1; l < L; l++) {
    const double v = c[(size_t)l * stride];
    if (v != v) { *has_nan = true; return l; }
    if (v < best) { best = v; idx = l; }
  }
  return idx;
}

// Open-loop line search (isls/isls.py:468-477): for every candidate alpha_l roll the model out from x^_0 with
// u^ + alpha_l du (rollout_batch, isls/isls.py:135-154), evaluate cost + sum((x-reg_x)^2 Qr) + sum((u-reg_u)^2 Rr),
// take the argmin unconditionally.  CTA = 32 problems x W warps, CPT candidates per thread (independent FP64
// chains, shared operand loads).  The control part of the cost comes from the per-problem quadratic (c0,c1,c2)
// accumulated by k_ff, so the hot loop is: 2 FMA for u, sincos, 5 FMA-type model updates (+ state terms).
// z <- clip(relax*x + (1-relax)*z + lam); r = x - z; lam += r   (isls/admm.py:43-59, projections.py:7-11).
// Written with explicit round-to-nearest intrinsics (no FMA contraction) so it is bit-identical to numpy on
// the same inputs.
__device__ __forceinline__ void admm_elem(double x, double relax, double lo, double hi, double &z, double &lam,
                                          double &rsq, double &dsq, int &mask) {
  const double pre = __dadd_rn(__dadd_rn(__dmul_rn(relax, x), __dmul_rn(__dsub_rn(1.0, relax), z)), lam);
  const double zn = fmin(fmax(pre, lo), hi);
  mask = (pre > hi) - (pre < lo);
  const double r = __dsub_rn(x, zn);
  const double dz = __dsub_rn(zn, z);
  lam = __dadd_rn(lam, r);
  z = zn;
  rsq = fma(r, r, rsq);
  dsq = fma(dz, dz, dsq);
}

template <class M>
__device__ __forceinline__ void admm_body(const Dev &d, const TileCtx<M> &c, int outer, int inner);
template <class M>
__device__ __forceinline__ void admm_finish(const Dev &d, const TileCtx<M> &c, int outer, int inner, int bi,
                                            double prim, double dual);

// ---- 1-D bulk async copies (TMA engine, cp.async.bulk) with mbarrier completion: the line search stages the next
// chunks of its per-step operands (u^, du, reg_x of the tile: contiguous in the tile-blocked layout) in shared memory
// while the FP64 chains run.  (A register software prefetch did not survive ptxas: it sank the copy of the prefetched
// registers to right behind the loads, exposing the full memory latency every step - 22 % of all stall samples sat
// on that one MOV, profiles/r1_linesearch_schedule.md.)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- ff-pass + linear rollout (same arithmetic as ff_body / k_ff_staged, bit for bit) with the per-step operands
// staged by 1-D TMA bulk copies: one warp per tile, lane 0 issues ONE cp.async.bulk per operand array and chunk of TC
// time steps (a tile's array is contiguous over time in the tile-blocked layout), NST chunks in flight, mbarrier
// completion.  Small batches run less than one warp per SM scheduler, so a launch lasts as long as ONE warp's
// instruction stream: against k_ff_staged (one 8-byte cp.async + address arithmetic per operand and thread, 250
// instructions per time step) this form issues a handful of copies per chunk from one lane, keeps up to NST * TC steps
// of operands in flight, reads plan constants from shared memory, and - with JC - takes the Jacobian scalars of the
// linearisation from the cache k_kpass wrote (d.Jc) instead of re-evaluating sincos in both sweeps.
// Finished lanes stay in the loop (the warp shares the copies) and are masked at the stores.
// Shared-memory carve-up of the TMA-staged ff-pass (one warp): [NST + 8 mbarriers][ring][rho_u][rho_x][qnz][seq]
template <class M, bool PX, bool JC, int TC, int NST>
struct FfTmaShape {
  static constexpr int n = M::n, m = M::m, nt = NTRI(M::m), NJ = M::NJA;
  static constexpr bool XB = !JC || PX;                   // the backward sweep stages x^ (Jacobian and / or cx)
  // backward slab: doubles per lane and step
  static constexpr int oJ = 0, oX = oJ + (JC ? NJ : 0), oU = oX + (XB ? n : 0), oQx = oU + m, oQu = oQx + m * n,
                       oQi = oQu + nt, oRu = oQi + nt, oRx = oRu + m, SB = oRx + (PX ? n : 0);
  // forward slab
  static constexpr int fK = 0, fk = fK + m * n, fU = fk + m, fRu = fU + m, fJ = fRu + m, SF = fJ + (JC ? NJ : n);
  static constexpr int NSTF_ = (NST * SB) / SF, NSTF = NSTF_ > 8 ? 8 : NSTF_;   // forward stages in the same memory
  static_assert(SF <= SB, "forward slab must fit the backward slab");
  static constexpr size_t RING_BYTES = (size_t)NST * TC * SB * TILE * sizeof(double);
  static size_t smem_bytes(int N) {
    return 128 + RING_BYTES + (size_t)N * (m + (PX ? n : 0)) * sizeof(double) + 2 * (size_t)N * sizeof(int);
  }
};

// One warp: plan constants into shared memory, mbarrier initialisation
template <class M, bool PX, bool JC, int TC, int NST>
__device__ __forceinline__ void ff_tma_setup(const Dev &d, double *smem_fft, const int lane, double *cst = nullptr,
                                             const bool fill_cst = true) {
  using S = FfTmaShape<M, PX, JC, TC, NST>;
  constexpr int n = S::n, m = S::m;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_fft);     // [NST] backward, [8] forward
  double *ring = smem_fft + 16;
  double *s_rhu = cst ? cst : ring + (size_t)NST * TC * S::SB * TILE;    // plan constants of all steps
  const int N = d.N;
  double *s_rhx = s_rhu + (size_t)N * m;
  int *s_qnz = reinterpret_cast<int *>(s_rhx + (PX ? (size_t)N * n : 0));
  int *s_seq = s_qnz + N;
  if (fill_cst) {
    for (int q = lane; q < N * m; q += TILE) s_rhu[q] = d.proj_u ? d.rho_u[q] : 0.0;
    if (PX)
      for (int q = lane; q < N * n; q += TILE) s_rhx[q] = d.rho_x[q];
    for (int q = lane; q < N; q += TILE) { s_qnz[q] = d.qnz[q]; s_seq[q] = d.seq[q]; }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NST + 8; i++) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
}

// One warp, one tile: ff-pass + linear rollout.  gb / gf = chunks consumed so far by this warp (the backward / forward
// rings keep running across tiles and calls).
template <class M, bool PX, bool JC, int TC, int NST>
__device__ __forceinline__ void ff_tma_tile(const Dev &d, double *smem_fft, const int tile, const int lane, unsigned &gb,
                                            unsigned &gf, double *cst = nullptr) {
  using S = FfTmaShape<M, PX, JC, TC, NST>;
  constexpr int n = S::n, m = S::m, nt = S::nt, NJ = S::NJ;
  constexpr bool XB = S::XB;
  constexpr int oJ = S::oJ, oX = S::oX, oU = S::oU, oQx = S::oQx, oQu = S::oQu, oQi = S::oQi, oRu = S::oRu, oRx = S::oRx,
                SB = S::SB;
  constexpr int fK = S::fK, fk = S::fk, fU = S::fU, fRu = S::fRu, fJ = S::fJ, SF = S::SF, NSTF = S::NSTF;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_fft);     // [NST] backward, [8] forward
  double *ring = smem_fft + 16;
  double *s_rhu = cst ? cst : ring + (size_t)NST * TC * SB * TILE;    // plan constants of all steps
  const int N = d.N;
  double *s_rhx = s_rhu + (size_t)N * m;
  int *s_qnz = reinterpret_cast<int *>(s_rhx + (PX ? (size_t)N * n : 0));
  int *s_seq = s_qnz + N;
  {
  TileCtx<M> c(d, tile, lane);
  const bool live = !(d.odone[c.b] || d.adone[c.b]);
  if (!__any_sync(0xffffffffu, live)) return;
  const double *xh = c.at(d.xh, d, n);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  double *kk = c.at(d.kk, d, m), *du = c.at(d.du, d, m);
  // tile base pointers of the staged arrays
  const size_t tb = (size_t)tile * N * TILE;
  const double *t_xh = d.xh + tb * n, *t_uh = d.uh + tb * m, *t_Qx = d.Qux + tb * (m * n), *t_Qu = d.Quu + tb * nt,
               *t_Qi = d.Qui + tb * nt, *t_ru = d.rgu + tb * m, *t_rx = PX ? d.rgx + tb * n : nullptr,
               *t_K = d.Kg + tb * (m * n), *t_kk = d.kk + tb * m, *t_J = JC ? d.Jc + tb * NJ : nullptr;
  const bool pu = d.proj_u != 0;
  // Rows of x^ the sweeps really need: the Jacobian reads the components >= M::JX0 only (car: theta, v), the quadratic
  // state cost reads all of them but only at via-point steps (plain loads there).  Without a state projection the rows
  // below JX0 are not staged at all: 44 -> 40 doubles per problem-step for the car.
  constexpr int JX0 = (!PX && !JC) ? M::JX0 : 0;
  const int xb0 = (JX0 > 0 && d.cost_kind == ISLS_COST_QUADRATIC) ? JX0 : 0;   // first staged row, backward sweep
  constexpr unsigned ROWB = TILE * sizeof(double);        // bytes of one component row of a step
  auto copy = [&](double *stage, int off, const double *src, int D, int t_lo, int cnt, unsigned long long *b) {
    bulk_g2s(stage + (size_t)off * TC * TILE, src + (size_t)t_lo * D * TILE, (unsigned)(cnt * D) * ROWB, b);
  };
  // ---- backward sweep: chunk ch covers t_hi = N-2 - ch*TC down to t_lo
  const int nchb = (N - 1 + TC - 1) / TC;
  auto issue_b = [&](int ch) {                            // lane 0 only
    const int st = (int)((gb + ch) % NST), t_hi = N - 2 - ch * TC, t_lo = max(0, t_hi - TC + 1), cnt = t_hi - t_lo + 1;
    double *sb = ring + (size_t)st * TC * SB * TILE;
    const int rows = (JC ? NJ : 0) + (XB ? n - xb0 : 0) + m + m * n + 2 * nt + (pu ? m : 0) + (PX ? n : 0);
    mbar_expect_tx(&bar[st], (unsigned)(cnt * rows) * ROWB);
    if (JC) copy(sb, oJ, t_J, NJ, t_lo, cnt, &bar[st]);
    if (XB) {
      if (xb0 == 0) copy(sb, oX, t_xh, n, t_lo, cnt, &bar[st]);
      else
        for (int tt = 0; tt < cnt; tt++)                    // rows xb0 .. n-1 of every step
          bulk_g2s(sb + ((size_t)oX * TC + tt * n + xb0) * TILE, t_xh + ((size_t)(t_lo + tt) * n + xb0) * TILE,
                   (unsigned)(n - xb0) * ROWB, &bar[st]);
    }
    copy(sb, oU, t_uh, m, t_lo, cnt, &bar[st]);
    copy(sb, oQx, t_Qx, m * n, t_lo, cnt, &bar[st]);
    copy(sb, oQu, t_Qu, nt, t_lo, cnt, &bar[st]);
    copy(sb, oQi, t_Qi, nt, t_lo, cnt, &bar[st]);
    if (pu) copy(sb, oRu, t_ru, m, t_lo, cnt, &bar[st]);
    if (PX) copy(sb, oRx, t_rx, n, t_lo, cnt, &bar[st]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NST; i++)
      if (i < nchb) issue_b(i);
  }
  double A[n][n], Bm[n][m];
  init_AB<M>(A, Bm);
  double v[n], rw2[m], rw[m];
#pragma unroll
  for (int j = 0; j < m; j++) { rw[j] = d.Rw[j]; rw2[j] = 2.0 * (d.u_std * rw[j]); }
  auto costgrad = [&](int t, const double (&x)[n], const double (&u)[m], const double (&rx)[n],
                      const double (&ru)[m], double (&cx)[n], double (&cu)[m], int qz, int sq,
                      const double (&rhx)[PX ? n : 1], const double (&rhu)[m]) {
    if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
      for (int i = 0; i < n; i++) cx[i] = qz ? ff_cx_quad(d.qd[t * n + i], x[i], EL(zs, n, sq, i)) : 0.0;
    } else {
      double ht[n];
      state_grad_hess<M>(d, zs, t, x, cx, ht);
    }
    if (PX) {
#pragma unroll
      for (int i = 0; i < n; i++) cx[i] = ff_pen(cx[i], rhx[PX ? i : 0], x[i], rx[i]);
    }
#pragma unroll
    for (int j = 0; j < m; j++) {
      double g = __dmul_rn(rw2[j], u[j]);
      if (pu) g = ff_pen(g, rhu[j], u[j], ru[j]);
      cu[j] = g;
    }
  };
  {   // terminal step N-1 (plain loads)
    double x[n], u[m], rx[n], ru[m], cx[n], cu[m];
#pragma unroll
    for (int i = 0; i < n; i++) { x[i] = EL(xh, n, N - 1, i); rx[i] = PX ? d.rgx[tb * n + ((size_t)(N - 1) * n + i) * TILE + lane] : 0.0; }
#pragma unroll
    for (int j = 0; j < m; j++) {
      u[j] = t_uh[((size_t)(N - 1) * m + j) * TILE + lane];
      ru[j] = pu ? t_ru[((size_t)(N - 1) * m + j) * TILE + lane] : 0.0;
    }
    double rhxT[PX ? n : 1], rhuT[m];
#pragma unroll
    for (int i = 0; i < (PX ? n : 1); i++) rhxT[i] = PX ? s_rhx[(N - 1) * n + i] : 0.0;
#pragma unroll
    for (int j = 0; j < m; j++) rhuT[j] = s_rhu[(N - 1) * m + j];
    costgrad(N - 1, x, u, rx, ru, cx, cu, s_qnz[N - 1], s_seq[N - 1], rhxT, rhuT);
#pragma unroll
    for (int i = 0; i < n; i++) v[i] = cx[i];
#pragma unroll
    for (int j = 0; j < m; j++) {
      const double cuu = __dmul_rn(2.0, __fma_rn(d.u_std, rw[j], d.rho_u[(N - 1) * m + j]));
      if (live) EL(kk, m, N - 1, j) = d.last_stage_dp ? 0.0 : -cu[j] / cuu;
    }
  }
  const bool need_x_glob = !XB;                           // x^ only where the state cost needs it (via-point steps)
  // The step's plan constants (rho, qnz, seq: warp-uniform shared-memory loads) are fetched ONE STEP AHEAD into
  // registers: read at their use inside costgrad each of them cost a uniform-address set-up (S2UR + ULEA) plus the
  // shared-memory latency right on the c_u -> q_u -> k chain of a lone warp (profiles/r2_tuning_log.md section 9).
  double rhx_n[PX ? n : 1], rhu_n[m];
  int qz_n, sq_n;
  auto fetch_consts = [&](int t) {
#pragma unroll
    for (int i = 0; i < (PX ? n : 1); i++) rhx_n[i] = PX ? s_rhx[t * n + i] : 0.0;
#pragma unroll
    for (int j = 0; j < m; j++) rhu_n[j] = s_rhu[t * m + j];
    qz_n = s_qnz[t];
    sq_n = s_seq[t];
  };
  fetch_consts(N >= 2 ? N - 2 : 0);
  for (int ch = 0; ch < nchb; ch++) {
    const int st = (int)((gb + ch) % NST), t_hi = N - 2 - ch * TC, t_lo = max(0, t_hi - TC + 1);
    const double *sb = ring + (size_t)st * TC * SB * TILE + lane;
    mbar_wait(&bar[st], (unsigned)(((gb + ch) / NST) & 1));
#pragma unroll 1
    for (int t = t_hi; t >= t_lo; t--) {
      const int tt = t - t_lo;
      double x[n], u[m], rx[n], ru[m], J[NJ], cx[n], cu[m], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
      const int qz = qz_n, sq = sq_n;
      double rhx_c[PX ? n : 1], rhu_c[m];
#pragma unroll
      for (int i = 0; i < (PX ? n : 1); i++) rhx_c[i] = rhx_n[i];
#pragma unroll
      for (int j = 0; j < m; j++) rhu_c[j] = rhu_n[j];
      if (t > 0) fetch_consts(t - 1);
#pragma unroll
      for (int i = 0; i < n; i++) {
        if (XB) x[i] = i >= xb0 ? sb[(size_t)(oX * TC + tt * n + i) * TILE] : (qz ? EL(xh, n, t, i) : 0.0);
        else x[i] = (need_x_glob && qz) ? EL(xh, n, t, i) : 0.0;
        rx[i] = PX ? sb[(size_t)(oRx * TC + tt * n + i) * TILE] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < m; j++) {
        u[j] = sb[(size_t)(oU * TC + tt * m + j) * TILE];
        ru[j] = pu ? sb[(size_t)(oRu * TC + tt * m + j) * TILE] : 0.0;
      }
#pragma unroll
      for (int a = 0; a < m; a++) {
#pragma unroll
        for (int j = 0; j < n; j++) Qux[a][j] = sb[(size_t)(oQx * TC + tt * (m * n) + a * n + j) * TILE];
#pragma unroll
        for (int b2 = 0; b2 <= a; b2++) {
          Quu[a][b2] = sb[(size_t)(oQu * TC + tt * nt + tri(a, b2)) * TILE]; Quu[b2][a] = Quu[a][b2];
          Qui[a][b2] = sb[(size_t)(oQi * TC + tt * nt + tri(a, b2)) * TILE]; Qui[b2][a] = Qui[a][b2];
        }
      }
      if (JC) {
#pragma unroll
        for (int q = 0; q < NJ; q++) J[q] = sb[(size_t)(oJ * TC + tt * NJ + q) * TILE];
      } else {
        M::jac(x, u, J, d.dt);
      }
      M::expand(J, A, Bm, d.dt);
      costgrad(t, x, u, rx, ru, cx, cu, qz, sq, rhx_c, rhu_c);
      ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
      if (live) {
#pragma unroll
        for (int j = 0; j < m; j++) EL(kk, m, t, j) = kt[j];
      }
    }
    __syncwarp();                                         // every lane is done with stage `st`
    if (lane == 0 && ch + NST < nchb) issue_b(ch + NST);
  }
  gb += (unsigned)nchb;
  // the forward sweep reads k through the async proxy: order this warp's generic-proxy stores of k before it
  fence_proxy_async();
  __syncwarp();
  // ---- forward sweep (linear rollout + control-cost polynomials): chunk ch covers t_lo = ch*TC upwards
  const int nchf = (N + TC - 1) / TC;
  unsigned long long *barf = bar + NST;
  auto issue_f = [&](int ch) {                            // lane 0 only
    const int st = (int)((gf + ch) % NSTF), t_lo = ch * TC, cnt = min(TC, N - t_lo);
    double *sf = ring + (size_t)st * TC * SF * TILE;
    const int rows = m * n + m + m + (pu ? m : 0) + (JC ? NJ : n - JX0);
    mbar_expect_tx(&barf[st], (unsigned)(cnt * rows) * ROWB);
    copy(sf, fK, t_K, m * n, t_lo, cnt, &barf[st]);
    copy(sf, fk, t_kk, m, t_lo, cnt, &barf[st]);
    copy(sf, fU, t_uh, m, t_lo, cnt, &barf[st]);
    if (pu) copy(sf, fRu, t_ru, m, t_lo, cnt, &barf[st]);
    if (JC) copy(sf, fJ, t_J, NJ, t_lo, cnt, &barf[st]);
    else if (JX0 == 0) copy(sf, fJ, t_xh, n, t_lo, cnt, &barf[st]);
    else
      for (int tt = 0; tt < cnt; tt++)
        bulk_g2s(sf + ((size_t)fJ * TC + tt * n + JX0) * TILE, t_xh + ((size_t)(t_lo + tt) * n + JX0) * TILE,
                 (unsigned)(n - JX0) * ROWB, &barf[st]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NSTF; i++)
      if (i < nchf) issue_f(i);
  }
  double dx[n];
#pragma unroll
  for (int i = 0; i < n; i++) dx[i] = 0.0;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
  for (int j = 0; j < m; j++) rhu_n[j] = s_rhu[j];         // step 0; every step fetches its successor's
  for (int ch = 0; ch < nchf; ch++) {
    const int st = (int)((gf + ch) % NSTF), t_lo = ch * TC, cnt = min(TC, N - t_lo);
    const double *sf = ring + (size_t)st * TC * SF * TILE + lane;
    mbar_wait(&barf[st], (unsigned)(((gf + ch) / NSTF) & 1));
#pragma unroll 1
    for (int tt = 0; tt < cnt; tt++) {
      const int t = t_lo + tt;
      double duv[m], u[m], K[m][n], kv[m], ru[m], rhu_c[m];
#pragma unroll
      for (int j = 0; j < m; j++) rhu_c[j] = rhu_n[j];
      if (t + 1 < N) {
#pragma unroll
        for (int j = 0; j < m; j++) rhu_n[j] = s_rhu[(t + 1) * m + j];
      }
#pragma unroll
      for (int a = 0; a < m; a++) {
#pragma unroll
        for (int j = 0; j < n; j++) K[a][j] = sf[(size_t)(fK * TC + tt * (m * n) + a * n + j) * TILE];
        kv[a] = sf[(size_t)(fk * TC + tt * m + a) * TILE];
        u[a] = sf[(size_t)(fU * TC + tt * m + a) * TILE];
        ru[a] = pu ? sf[(size_t)(fRu * TC + tt * m + a) * TILE] : 0.0;
      }
#pragma unroll
      for (int a = 0; a < m; a++) {
        double acc = 0.0;
        if (t < N - 1) {
#pragma unroll
          for (int j = 0; j < n; j++) acc = fma(K[a][j], dx[j], acc);
        }
        duv[a] = acc + kv[a];
        if (live) EL(du, m, t, a) = duv[a];
        r0 = fma(rw[a] * u[a], u[a], r0);
        r1 = fma(rw[a] * u[a], duv[a], r1);
        r2 = fma(rw[a] * duv[a], duv[a], r2);
        if (pu) {
          const double rho = rhu_c[a], e = u[a] - ru[a];
          c0 = fma(rho * e, e, c0);
          c1 = fma(2.0 * rho * e, duv[a], c1);
          c2 = fma(rho * duv[a], duv[a], c2);
        }
      }
      if (t < N - 1) {
        double J[NJ], dxn[n];
        if (JC) {
#pragma unroll
          for (int q = 0; q < NJ; q++) J[q] = sf[(size_t)(fJ * TC + tt * NJ + q) * TILE];
        } else {
          double x[n];
#pragma unroll
          for (int i = 0; i < n; i++) x[i] = i >= JX0 ? sf[(size_t)(fJ * TC + tt * n + i) * TILE] : 0.0;   // jac reads rows >= JX0
          M::jac(x, u, J, d.dt);
        }
        M::expand(J, A, Bm, d.dt);
        mat_Ax_Bu<M>(A, Bm, dx, duv, dxn);
#pragma unroll
        for (int i = 0; i < n; i++) dx[i] = dxn[i];
      }
    }
    __syncwarp();
    if (lane == 0 && ch + NSTF < nchf) issue_f(ch + NSTF);
  }
  if (live) {
    const size_t S = (size_t)d.T * TILE;
    r0 *= d.u_std;
    r1 *= 2.0 * d.u_std;
    r2 *= d.u_std;
    d.cq[c.b] = c0 + r0;
    d.cq[S + c.b] = c1 + r1;
    d.cq[2 * S + c.b] = c2 + r2;
    d.cq[3 * S + c.b] = r0;
    d.cq[4 * S + c.b] = r1;
    d.cq[5 * S + c.b] = r2;
  }
  gf += (unsigned)nchf;
  __syncwarp();                 // the ring is re-used by the next tile's backward sweep
  }
}

template <class M, bool PX, bool JC, int TC, int NST>
__global__ void __launch_bounds__(TILE) k_ff_tma(Dev d) {
  extern __shared__ __align__(128) double smem_fft[];
  const int lane = threadIdx.x;
  const int ntiles = (d.tile1 - d.tile0 + d.tstep - 1) / d.tstep;
  if ((int)blockIdx.x >= ntiles) return;
  ff_tma_setup<M, PX, JC, TC, NST>(d, smem_fft, lane);
  unsigned gb = 0, gf = 0;
  // persistent over tiles (grid = tiles for the small-batch launches, fewer CTAs in the overlapped schedule)
  for (int it = blockIdx.x; it < ntiles; it += gridDim.x)
    ff_tma_tile<M, PX, JC, TC, NST>(d, smem_fft, d.tile0 + it * d.tstep, lane, gb, gf);
}

// ---- Warp-split form of k_ff_tma for models with larger state (arm: n = 9, m = 3) at small batches: G warps share one
// tile.  Lane = problem as everywhere else (same tile-blocked layout, same TMA-staged operand ring), but the rows of the
// n-sized algebra are split over the warps - warp w owns the components i with i mod G = w of v, qx, cx, dx - and the
// recursion state (v in the backward sweep, dx in the forward one) is exchanged through shared memory with one CTA
// barrier per time step.  The m-sized algebra (qu, k, the refinement step) is computed by every warp (it is cheaper than
// a second exchange).  Every output element is accumulated in the same order as in ff_step / mat_Ax_Bu, so the results
// are bit-identical to the one-thread-per-trajectory kernels; what changes is that a lone warp's 500-instruction step
// becomes four ~120-instruction streams on four SM sub-partitions.  (north_star: "one warp or thread-block per
// trajectory ... the small dense algebra split across it"; here a thread block per TILE of trajectories, so the loads
// stay full 256-byte lines.)
template <class M, bool PX, bool JC, int G, int TC, int NST>
__global__ void __launch_bounds__(TILE * G, 4) k_ff_ws(Dev d) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m), NJ = M::NJA;
  constexpr bool XB = !JC || PX;
  constexpr int oJ = 0, oX = oJ + (JC ? NJ : 0), oU = oX + (XB ? n : 0), oQx = oU + m, oQu = oQx + m * n,
                oQi = oQu + nt, oRu = oQi + nt, oRx = oRu + m, SB = oRx + (PX ? n : 0);
  constexpr int fK = 0, fk = fK + m * n, fU = fk + m, fRu = fU + m, fJ = fRu + m, SF = fJ + (JC ? NJ : n);
  constexpr int NSTF_ = (NST * SB) / SF, NSTF = NSTF_ > 8 ? 8 : NSTF_;
  static_assert(SF <= SB, "forward slab must fit the backward slab");
  extern __shared__ __align__(128) double smem_ffw[];
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_ffw);     // [NST] backward, [8] forward
  double *ring = smem_ffw + 16;
  double *xch = ring + (size_t)NST * TC * SB * TILE;      // [2][n][TILE] exchange buffer of the recursion state
  double *s_rhu = xch + 2 * n * TILE;
  const int N = d.N;
  double *s_rhx = s_rhu + (size_t)N * m;
  int *s_qnz = reinterpret_cast<int *>(s_rhx + (PX ? (size_t)N * n : 0));
  int *s_seq = s_qnz + N;
  const int lane = threadIdx.x, w = threadIdx.y, tid = w * TILE + lane;
  const bool leader = tid == 0;
  const int ntiles = (d.tile1 - d.tile0 + d.tstep - 1) / d.tstep;
  if ((int)blockIdx.x >= ntiles) return;
  for (int q = tid; q < N * m; q += TILE * G) s_rhu[q] = d.proj_u ? d.rho_u[q] : 0.0;
  if (PX)
    for (int q = tid; q < N * n; q += TILE * G) s_rhx[q] = d.rho_x[q];
  for (int q = tid; q < N; q += TILE * G) { s_qnz[q] = d.qnz[q]; s_seq[q] = d.seq[q]; }
  if (leader) {
#pragma unroll
    for (int i = 0; i < NST + 8; i++) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned gb = 0, gf = 0;
  auto own = [&](int i) -> bool { return (i % G) == w; };
  for (int it = blockIdx.x; it < ntiles; it += gridDim.x) {
    const int tile = d.tile0 + it * d.tstep;
    TileCtx<M> c(d, tile, lane);
    const bool live = !(d.odone[c.b] || d.adone[c.b]);
    if (!__syncthreads_or(live)) continue;
    const double *xh = c.at(d.xh, d, n);
    const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
    double *kk = c.at(d.kk, d, m), *du = c.at(d.du, d, m);
    const size_t tb = (size_t)tile * N * TILE;
    const double *t_xh = d.xh + tb * n, *t_uh = d.uh + tb * m, *t_Qx = d.Qux + tb * (m * n), *t_Qu = d.Quu + tb * nt,
                 *t_Qi = d.Qui + tb * nt, *t_ru = d.rgu + tb * m, *t_rx = PX ? d.rgx + tb * n : nullptr,
                 *t_K = d.Kg + tb * (m * n), *t_kk = d.kk + tb * m, *t_J = JC ? d.Jc + tb * NJ : nullptr;
    const bool pu = d.proj_u != 0;
    constexpr unsigned ROWB = TILE * sizeof(double);
    auto copy = [&](double *stage, int off, const double *src, int D, int t_lo, int cnt, unsigned long long *b) {
      bulk_g2s(stage + (size_t)off * TC * TILE, src + (size_t)t_lo * D * TILE, (unsigned)(cnt * D) * ROWB, b);
    };
    const int nchb = (N - 1 + TC - 1) / TC;
    auto issue_b = [&](int ch) {                            // leader only
      const int st = (int)((gb + ch) % NST), t_hi = N - 2 - ch * TC, t_lo = max(0, t_hi - TC + 1), cnt = t_hi - t_lo + 1;
      double *sb = ring + (size_t)st * TC * SB * TILE;
      const int rows = (JC ? NJ : 0) + (XB ? n : 0) + m + m * n + 2 * nt + (pu ? m : 0) + (PX ? n : 0);
      mbar_expect_tx(&bar[st], (unsigned)(cnt * rows) * ROWB);
      if (JC) copy(sb, oJ, t_J, NJ, t_lo, cnt, &bar[st]);
      if (XB) copy(sb, oX, t_xh, n, t_lo, cnt, &bar[st]);
      copy(sb, oU, t_uh, m, t_lo, cnt, &bar[st]);
      copy(sb, oQx, t_Qx, m * n, t_lo, cnt, &bar[st]);
      copy(sb, oQu, t_Qu, nt, t_lo, cnt, &bar[st]);
      copy(sb, oQi, t_Qi, nt, t_lo, cnt, &bar[st]);
      if (pu) copy(sb, oRu, t_ru, m, t_lo, cnt, &bar[st]);
      if (PX) copy(sb, oRx, t_rx, n, t_lo, cnt, &bar[st]);
    };
    if (leader) {
#pragma unroll
      for (int i = 0; i < NST; i++)
        if (i < nchb) issue_b(i);
    }
    double A[n][n], Bm[n][m];
    init_AB<M>(A, Bm);
    double v[n], rw2[m], rw[m];
#pragma unroll
    for (int j = 0; j < m; j++) { rw[j] = d.Rw[j]; rw2[j] = 2.0 * (d.u_std * rw[j]); }
    // cx_i (one component) and cu (all m) with the arithmetic of the other ff variants
    auto cx_of = [&](int t, int i, double xi, double rxi, int qz, int sq) -> double {
      double cxi = qz ? ff_cx_quad(d.qd[t * n + i], xi, EL(zs, n, sq, i)) : 0.0;
      if (PX) cxi = ff_pen(cxi, s_rhx[t * n + i], xi, rxi);
      return cxi;
    };
    auto cu_of = [&](int t, const double (&u)[m], const double (&ru)[m], double (&cu)[m]) {
#pragma unroll
      for (int j = 0; j < m; j++) {
        double g = __dmul_rn(rw2[j], u[j]);
        if (pu) g = ff_pen(g, s_rhu[t * m + j], u[j], ru[j]);
        cu[j] = g;
      }
    };
    {   // terminal step N-1: every warp computes the whole v (plain loads; once per launch)
      double u[m], ru[m], cu[m];
      const int qz = s_qnz[N - 1], sq = s_seq[N - 1];
#pragma unroll
      for (int i = 0; i < n; i++) {
        const double xi = EL(xh, n, N - 1, i), rxi = PX ? d.rgx[tb * n + ((size_t)(N - 1) * n + i) * TILE + lane] : 0.0;
        v[i] = cx_of(N - 1, i, xi, rxi, qz, sq);
      }
#pragma unroll
      for (int j = 0; j < m; j++) {
        u[j] = t_uh[((size_t)(N - 1) * m + j) * TILE + lane];
        ru[j] = pu ? t_ru[((size_t)(N - 1) * m + j) * TILE + lane] : 0.0;
      }
      cu_of(N - 1, u, ru, cu);
#pragma unroll
      for (int j = 0; j < m; j++) {
        const double cuu = __dmul_rn(2.0, __fma_rn(d.u_std, rw[j], d.rho_u[(N - 1) * m + j]));
        if (live && w == 0) EL(kk, m, N - 1, j) = d.last_stage_dp ? 0.0 : -cu[j] / cuu;
      }
    }
    int xb = 0;                                             // exchange buffer of this step
    for (int ch = 0; ch < nchb; ch++) {
      const int st = (int)((gb + ch) % NST), t_hi = N - 2 - ch * TC, t_lo = max(0, t_hi - TC + 1);
      const double *sb = ring + (size_t)st * TC * SB * TILE + lane;
      mbar_wait(&bar[st], (unsigned)(((gb + ch) / NST) & 1));
#pragma unroll 1
      for (int t = t_hi; t >= t_lo; t--) {
        const int tt = t - t_lo;
        double u[m], ru[m], J[NJ], cu[m], Quu[m][m], Qui[m][m], qu[m], k0[m], wv[m], kt[m];
        const int qz = s_qnz[t], sq = s_seq[t];
#pragma unroll
        for (int j = 0; j < m; j++) {
          u[j] = sb[(size_t)(oU * TC + tt * m + j) * TILE];
          ru[j] = pu ? sb[(size_t)(oRu * TC + tt * m + j) * TILE] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < m; a++)
#pragma unroll
          for (int b2 = 0; b2 <= a; b2++) {
            Quu[a][b2] = sb[(size_t)(oQu * TC + tt * nt + tri(a, b2)) * TILE]; Quu[b2][a] = Quu[a][b2];
            Qui[a][b2] = sb[(size_t)(oQi * TC + tt * nt + tri(a, b2)) * TILE]; Qui[b2][a] = Qui[a][b2];
          }
        if (JC) {
#pragma unroll
          for (int q = 0; q < NJ; q++) J[q] = sb[(size_t)(oJ * TC + tt * NJ + q) * TILE];
        } else {
          double x[n];
#pragma unroll
          for (int i = 0; i < n; i++) x[i] = sb[(size_t)(oX * TC + tt * n + i) * TILE];
          M::jac(x, u, J, d.dt);
        }
        M::expand(J, A, Bm, d.dt);
        cu_of(t, u, ru, cu);
        // ff_step, m-sized part on every warp: qu = cu + B'v, k0 = -Quu^-1 qu, w = qu + Quu k0, k = k0 - Quu^-1 w
        mat_Bt_v<M>(Bm, v, qu);
#pragma unroll
        for (int j = 0; j < m; j++) qu[j] += cu[j];
#pragma unroll
        for (int a = 0; a < m; a++) {
          double acc = 0.0;
#pragma unroll
          for (int b2 = 0; b2 < m; b2++) acc = fma(Qui[a][b2], qu[b2], acc);
          k0[a] = -acc;
        }
#pragma unroll
        for (int a = 0; a < m; a++) {
          double acc = qu[a];
#pragma unroll
          for (int b2 = 0; b2 < m; b2++) acc = fma(Quu[a][b2], k0[b2], acc);
          wv[a] = acc;
        }
#pragma unroll
        for (int a = 0; a < m; a++) {
          double acc = 0.0;
#pragma unroll
          for (int b2 = 0; b2 < m; b2++) acc = fma(Qui[a][b2], wv[b2], acc);
          kt[a] = k0[a] - acc;
        }
        if (live && w == 0) {
#pragma unroll
          for (int j = 0; j < m; j++) EL(kk, m, t, j) = kt[j];
        }
        // n-sized part, own rows: qx_i = cx_i + (A'v)_i, v_i <- qx_i + sum_a Qux[a][i] k[a]
        double *xo = xch + (size_t)xb * n * TILE + lane;
#pragma unroll
        for (int i = 0; i < n; i++) {
          if (own(i)) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < n; k++) {
              if (M::am(k, i) == MO) acc += v[k];
              else if (M::am(k, i) == MV) acc = fma(A[k][i], v[k], acc);
            }
            double xi = 0.0, rxi = 0.0;
            if (XB) xi = sb[(size_t)(oX * TC + tt * n + i) * TILE];
            else if (qz) xi = EL(xh, n, t, i);
            if (PX) rxi = sb[(size_t)(oRx * TC + tt * n + i) * TILE];
            double vi = acc + cx_of(t, i, xi, rxi, qz, sq);
#pragma unroll
            for (int a = 0; a < m; a++) vi = fma(sb[(size_t)(oQx * TC + tt * (m * n) + a * n + i) * TILE], kt[a], vi);
            xo[(size_t)i * TILE] = vi;
          }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < n; i++) v[i] = xo[(size_t)i * TILE];
        xb ^= 1;
      }
      // the per-step barrier above also means every warp is done with stage `st`
      if (leader && ch + NST < nchb) issue_b(ch + NST);
    }
    gb += (unsigned)nchb;
    fence_proxy_async();                                    // k (written above by warp 0) is read through the async proxy below
    __syncthreads();
    // ---- forward sweep
    const int nchf = (N + TC - 1) / TC;
    unsigned long long *barf = bar + NST;
    auto issue_f = [&](int ch) {                            // leader only
      const int st = (int)((gf + ch) % NSTF), t_lo = ch * TC, cnt = min(TC, N - t_lo);
      double *sf = ring + (size_t)st * TC * SF * TILE;
      const int rows = m * n + m + m + (pu ? m : 0) + (JC ? NJ : n);
      mbar_expect_tx(&barf[st], (unsigned)(cnt * rows) * ROWB);
      copy(sf, fK, t_K, m * n, t_lo, cnt, &barf[st]);
      copy(sf, fk, t_kk, m, t_lo, cnt, &barf[st]);
      copy(sf, fU, t_uh, m, t_lo, cnt, &barf[st]);
      if (pu) copy(sf, fRu, t_ru, m, t_lo, cnt, &barf[st]);
      if (JC) copy(sf, fJ, t_J, NJ, t_lo, cnt, &barf[st]);
      else copy(sf, fJ, t_xh, n, t_lo, cnt, &barf[st]);
    };
    if (leader) {
#pragma unroll
      for (int i = 0; i < NSTF; i++)
        if (i < nchf) issue_f(i);
    }
    double dx[n];
#pragma unroll
    for (int i = 0; i < n; i++) dx[i] = 0.0;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
    const bool polys = w == G - 1;                          // the warp with the fewest rows accumulates the cost polynomials
    for (int ch = 0; ch < nchf; ch++) {
      const int st = (int)((gf + ch) % NSTF), t_lo = ch * TC, cnt = min(TC, N - t_lo);
      const double *sf = ring + (size_t)st * TC * SF * TILE + lane;
      mbar_wait(&barf[st], (unsigned)(((gf + ch) / NSTF) & 1));
#pragma unroll 1
      for (int tt = 0; tt < cnt; tt++) {
        const int t = t_lo + tt;
        double duv[m], u[m];
#pragma unroll
        for (int a = 0; a < m; a++) {
          u[a] = sf[(size_t)(fU * TC + tt * m + a) * TILE];
          double acc = 0.0;
          if (t < N - 1) {
#pragma unroll
            for (int j = 0; j < n; j++) acc = fma(sf[(size_t)(fK * TC + tt * (m * n) + a * n + j) * TILE], dx[j], acc);
          }
          duv[a] = acc + sf[(size_t)(fk * TC + tt * m + a) * TILE];
        }
        if (live && w == 0) {
#pragma unroll
          for (int a = 0; a < m; a++) EL(du, m, t, a) = duv[a];
        }
        if (polys) {
#pragma unroll
          for (int a = 0; a < m; a++) {
            r0 = fma(rw[a] * u[a], u[a], r0);
            r1 = fma(rw[a] * u[a], duv[a], r1);
            r2 = fma(rw[a] * duv[a], duv[a], r2);
            if (pu) {
              const double rho = s_rhu[t * m + a], e = u[a] - sf[(size_t)(fRu * TC + tt * m + a) * TILE];
              c0 = fma(rho * e, e, c0);
              c1 = fma(2.0 * rho * e, duv[a], c1);
              c2 = fma(rho * duv[a], duv[a], c2);
            }
          }
        }
        if (t < N - 1) {
          double J[NJ];
          if (JC) {
#pragma unroll
            for (int q = 0; q < NJ; q++) J[q] = sf[(size_t)(fJ * TC + tt * NJ + q) * TILE];
          } else {
            double x[n];
#pragma unroll
            for (int i = 0; i < n; i++) x[i] = sf[(size_t)(fJ * TC + tt * n + i) * TILE];
            M::jac(x, u, J, d.dt);
          }
          M::expand(J, A, Bm, d.dt);
          double *xo = xch + (size_t)xb * n * TILE + lane;
#pragma unroll
          for (int i = 0; i < n; i++) {
            if (own(i)) {                                  // row i of mat_Ax_Bu
              double acc = 0.0;
#pragma unroll
              for (int k = 0; k < n; k++) {
                if (M::am(i, k) == MO) acc += dx[k];
                else if (M::am(i, k) == MV) acc = fma(A[i][k], dx[k], acc);
              }
#pragma unroll
              for (int k = 0; k < m; k++) {
                if (M::bm(i, k) == MO) acc += duv[k];
                else if (M::bm(i, k) == MV) acc = fma(Bm[i][k], duv[k], acc);
              }
              xo[(size_t)i * TILE] = acc;
            }
          }
          __syncthreads();
#pragma unroll
          for (int i = 0; i < n; i++) dx[i] = xo[(size_t)i * TILE];
          xb ^= 1;
        }
      }
      __syncthreads();                                      // (the last step of the sweep has no exchange barrier)
      if (leader && ch + NSTF < nchf) issue_f(ch + NSTF);
    }
    if (live && polys) {
      const size_t S = (size_t)d.T * TILE;
      r0 *= d.u_std;
      r1 *= 2.0 * d.u_std;
      r2 *= d.u_std;
      d.cq[c.b] = c0 + r0;
      d.cq[S + c.b] = c1 + r1;
      d.cq[2 * S + c.b] = c2 + r2;
      d.cq[3 * S + c.b] = r0;
      d.cq[4 * S + c.b] = r1;
      d.cq[5 * S + c.b] = r2;
    }
    gf += (unsigned)nchf;
    __syncthreads();
  }
}

// `fuse` != 0: the CTA finishes with the streaming ADMM z / lambda / reg update of the winner (control-only
// projections).  PX = the plan has a state projection (compile-time, so the control-only kernel carries no penalty
// accumulators or reg_x operands).  Every thread of the CTA runs the rollout loop (it contains CTA barriers); lanes
// whose problem is finished compute on stale data and are masked at the writes.
// One tile's line search; returns the number of staged chunks it consumed (the mbarrier ring keeps running across the
// tiles of a persistent CTA: `gch` = chunks consumed so far by this CTA, `init_bars` = first tile of the CTA).
template <class M, int CPT, int MAXW, bool PX, int TCMAX = 10>
__device__ __forceinline__ int ls_tile(const Dev &d, const int tile, int fuse, int outer, int inner, const unsigned gch,
                                       const bool init_bars) {
  constexpr int n = M::n, m = M::m;
  constexpr int LMAX = CPT * MAXW;                            // candidates this CTA shape can hold (>= d.L)
  constexpr int ROWS = 2 * m + (PX ? n : 0);                  // staged doubles per lane and step
  constexpr int STAGE_BYTES = LMAX > 32 ? 4096 : 12288;       // keeps the static shared memory under 48 KB
  constexpr int TC_ = STAGE_BYTES / (ROWS * TILE * 8);        // steps per chunk
  constexpr int TC = TC_ > TCMAX ? TCMAX : (TC_ < 1 ? 1 : TC_);
  constexpr int NST = 2;                                      // chunks in flight
  __shared__ __align__(128) double su[NST][TC][m][TILE];      // u^
  __shared__ __align__(128) double sd[NST][TC][m][TILE];      // du
  __shared__ __align__(128) double sr[NST][PX ? TC : 1][PX ? n : 1][TILE];   // reg_x
  __shared__ unsigned long long fullbar[NST];
  __shared__ double sc[LMAX][TILE];
  __shared__ double scs[LMAX][TILE];         // state-cost part of every candidate (cost of the winner w/o penalty)
  __shared__ double sred[2][MAXW][TILE];     // residual partial sums of the fused ADMM update
  __shared__ int sbest[TILE];
  TileCtx<M> c(d, tile, threadIdx.x);
  const int w = threadIdx.y;
  const bool skip = d.odone[c.b] || d.adone[c.b];
  if (__syncthreads_and(skip)) return 0;
  const int nchunks = (d.N + TC - 1) / TC;
  {
    const bool leader = threadIdx.x == 0 && w == 0;
    const double *uh_t = d.uh + (size_t)tile * d.N * m * TILE, *du_t = d.du + (size_t)tile * d.N * m * TILE;
    const double *rg_t = PX ? d.rgx + (size_t)tile * d.N * n * TILE : nullptr;
    auto issue = [&](int ch) {                                // leader only
      const int st = (int)((gch + ch) % NST), t0 = ch * TC, cnt = min(TC, d.N - t0);
      const unsigned bu = (unsigned)(cnt * m * TILE * sizeof(double)), bx = (unsigned)(cnt * n * TILE * sizeof(double));
      mbar_expect_tx(&fullbar[st], 2 * bu + (PX ? bx : 0u));
      bulk_g2s(&su[st][0][0][0], uh_t + (size_t)t0 * m * TILE, bu, &fullbar[st]);
      bulk_g2s(&sd[st][0][0][0], du_t + (size_t)t0 * m * TILE, bu, &fullbar[st]);
      if (PX) bulk_g2s(&sr[st][0][0][0], rg_t + (size_t)t0 * n * TILE, bx, &fullbar[st]);
    };
    if (leader && init_bars) {
#pragma unroll
      for (int i = 0; i < NST; i++) mbar_init(&fullbar[i], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (leader) {
#pragma unroll
      for (int i = 0; i < NST; i++)
        if (i < nchunks) issue(i);
    }
    const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m), *du = c.at(d.du, d, m);
    const double *rgx = c.at(d.rgx, d, n);
    const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
    double al[CPT], x[CPT][n], cs[CPT], px[CPT];
    bool bad[CPT];
#pragma unroll
    for (int q = 0; q < CPT; q++) {
      const int l = w * CPT + q;
      al[q] = l < d.L ? d.alphas[l] : 0.0;
      cs[q] = px[q] = 0.0;
#pragma unroll
      for (int i = 0; i < n; i++) x[q][i] = EL(xh, n, 0, i);
      bad[q] = !M::fast_state(x[q]);
    }
    for (int ch = 0; ch < nchunks; ch++) {
      const int st = (int)((gch + ch) % NST), t0 = ch * TC, cnt = min(TC, d.N - t0);
      unsigned qmask = 0;                                     // steps of this chunk that carry a state cost
#pragma unroll
      for (int tt = 0; tt < TC; tt++)
        if (tt < cnt && d.qnz[t0 + tt]) qmask |= 1u << tt;
      mbar_wait(&fullbar[st], (unsigned)(((gch + ch) / NST) & 1));
#pragma unroll 1   // (unroll 2 made the loop 48 KB of SASS, past the 32 KB L1.5 instruction cache: 7 % no_instruction stalls; 74.0 -> 72.5 us at 8,192 problems)
      for (int tt = 0; tt < cnt; tt++) {
        const int t = t0 + tt;
        double un[m], dun[m], rx[n], zv[n], qd[n], rhx[n];
        const bool qz = (qmask >> tt) & 1u;
#pragma unroll
        for (int j = 0; j < m; j++) { un[j] = su[st][tt][j][c.lane]; dun[j] = sd[st][tt][j][c.lane]; }
        if (PX) {
#pragma unroll
          for (int i = 0; i < n; i++) { rx[i] = sr[st][tt][i][c.lane]; rhx[i] = d.rho_x[t * n + i]; }
        }
        if (qz) {
          const int s = d.seq[t];
#pragma unroll
          for (int i = 0; i < n; i++) { zv[i] = EL(zs, n, s, i); qd[i] = d.qd[t * n + i]; }
          if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
            for (int q = 0; q < CPT; q++)
#pragma unroll
              for (int i = 0; i < n; i++) { const double e = x[q][i] - zv[i]; cs[q] += (e * e) * qd[i]; }
          } else {
#pragma unroll
            for (int q = 0; q < CPT; q++) cs[q] += state_cost<M>(d, zs, t, x[q]);
          }
        }
        if (PX) {
#pragma unroll
          for (int q = 0; q < CPT; q++)
#pragma unroll
            for (int i = 0; i < n; i++) { const double e = x[q][i] - rx[i]; px[q] += (e * e) * rhx[i]; }
        }
        // branch-free model steps, written stage by stage across the CPT chains (models.cuh: steps_fast); an argument
        // outside the fast range of sincos / mod only raises the candidate's sticky flag (its rollout is redone below)
        double u[CPT][m];
#pragma unroll
        for (int j = 0; j < m; j++)
#pragma unroll
          for (int q = 0; q < CPT; q++) u[q][j] = fma(al[q], dun[j], un[j]);
        M::template steps_fast<CPT>(x, u, d.dt, bad);
      }
      __syncthreads();                                        // every warp is done with stage `st`
      if (leader && ch + NST < nchunks) issue(ch + NST);
    }
#pragma unroll
    for (int q = 0; q < CPT; q++) {
      if (bad[q] && !skip) {   // rare: exact re-rollout of this candidate (library sincos / fmod paths)
        double xe[n], u[m], xn[n];
        double cse = 0.0, pxe = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) xe[i] = EL(xh, n, 0, i);
        for (int t = 0; t < d.N; t++) {
#pragma unroll
          for (int j = 0; j < m; j++) u[j] = fma(al[q], EL(du, m, t, j), EL(uh, m, t, j));
          cse += state_cost<M>(d, zs, t, xe);
          if (PX) {
#pragma unroll
            for (int i = 0; i < n; i++) { const double e = xe[i] - EL(rgx, n, t, i); pxe += (e * e) * d.rho_x[t * n + i]; }
          }
          M::step(xe, u, xn, d.dt);
#pragma unroll
          for (int i = 0; i < n; i++) xe[i] = xn[i];
        }
        cs[q] = cse;
        px[q] = pxe;
      }
    }
    const size_t S = (size_t)d.T * TILE;
    // isls_admm searches on the cost alone (isls.py:586-599): R-only polynomial (rows 3-5)
    const int pr0 = d.ls_cost_only ? 3 : 0;
    const double c0 = d.cq[pr0 * S + c.b], c1 = d.cq[(pr0 + 1) * S + c.b], c2 = d.cq[(pr0 + 2) * S + c.b];
#pragma unroll
    for (int q = 0; q < CPT; q++) {
      const int l = w * CPT + q;
      if (l < d.L && !skip) {
        double tot = cs[q] + fma(al[q], fma(al[q], c2, c1), c0);   // cost_function + control penalty (isls.py:470,476)
        if (d.cost_kind != ISLS_COST_QUADRATIC && tot != tot) tot = 1e6;   // Tutorial cell 14: cost closure, NaN -> 1e6
        if (PX && !d.ls_cost_only) tot += px[q];                   // isls.py:473 (isls_admm: no penalty, isls.py:588-590)
        sc[l][c.lane] = tot;
        scs[l][c.lane] = cs[q];
      }
    }
  }
  __syncthreads();
  if (w == 0 && !skip) {
    bool has_nan;
    const int idx = argmin_np(&sc[0][c.lane], d.L, TILE, &has_nan);
    d.best[c.b] = idx;
    d.best_cost[c.b] = sc[idx][c.lane];
    sbest[c.lane] = idx;
    if (has_nan) d.status[c.b] |= ISLS_ST_NAN_COST;
    if (d.lsc) {
      double *o = d.lsc + (size_t)tile * d.L * TILE + c.lane;
      for (int l = 0; l < d.L; l++) o[(size_t)l * TILE] = sc[l][c.lane];
    }
  }
  if (!fuse) return nchunks;
  // ---- fused ADMM update for control-only projections (admm.py:43-97): z_u, lambda_u and reg_u depend on the winner
  // only through u = u^ + alpha* du, so the whole CTA streams the tile's N x m elements (warp w takes t = w, w+W, ..)
  // instead of a separate trajectory-per-thread kernel re-rolling the model.  Residual sums: fixed-order reduction.
  __syncthreads();
  const int W = blockDim.y;
  double pru = 0.0, dru = 0.0;
  if (!skip) {
    const double *uh = c.at(d.uh, d, m), *du = c.at(d.du, d, m);
    double *zu = c.at(d.zu, d, m), *lu = c.at(d.lu, d, m), *rgu = c.at(d.rgu, d, m);
    const double al = d.alphas[sbest[c.lane]];
    int8_t *mku = (d.out.mask_u && c.valid) ? d.out.mask_u + c.ob * d.N * m : nullptr;
    // UB time steps in flight per thread.  The load phase is branch-free (t clamped to N-1, only the stores are
    // predicated) and holds the raw operands: with the guards and u = u^ + alpha du inside it, ptxas emitted UB separate
    // load-then-use blocks, i.e. UB serialised memory round trips per batch (28 per tile at N = 100: a fifth of all stall
    // samples at 8,192 problems, profiles/r2_tuning_log.md section 8).
    constexpr int UB = 4;
    for (int t0 = w; t0 < d.N; t0 += W * UB) {
      double dv[UB][m], hv[UB][m], zv[UB][m], lv[UB][m], lo[UB][m], hi[UB][m];
#pragma unroll
      for (int q = 0; q < UB; q++) {
        const int t = min(t0 + q * W, d.N - 1);
#pragma unroll
        for (int j = 0; j < m; j++) {
          dv[q][j] = EL(du, m, t, j);
          hv[q][j] = EL(uh, m, t, j);
          zv[q][j] = EL(zu, m, t, j);
          lv[q][j] = EL(lu, m, t, j);
          lo[q][j] = d.lo_u[t * m + j];
          hi[q][j] = d.hi_u[t * m + j];
        }
      }
#pragma unroll
      for (int q = 0; q < UB; q++) {
        const int t = t0 + q * W;
        if (t < d.N) {
#pragma unroll
          for (int j = 0; j < m; j++) {
            int mk;
            const double uv = fma(al, dv[q][j], hv[q][j]);                // identical to the candidate's u
            admm_elem(uv, d.relax, lo[q][j], hi[q][j], zv[q][j], lv[q][j], pru, dru, mk);
            EL(zu, m, t, j) = zv[q][j];
            EL(lu, m, t, j) = lv[q][j];
            EL(rgu, m, t, j) = __dsub_rn(zv[q][j], lv[q][j]);
            if (mku) mku[t * m + j] = (int8_t)mk;
          }
        }
      }
    }
  }
  sred[0][w][c.lane] = pru;
  sred[1][w][c.lane] = dru;
  __syncthreads();
  if (w == 0 && !skip) {
    double ps = 0.0, ds = 0.0;
    for (int q = 0; q < W; q++) { ps += sred[0][q][c.lane]; ds += sred[1][q][c.lane]; }
    const int bi = sbest[c.lane];
    const double al = d.alphas[bi];
    const size_t S = (size_t)d.T * TILE;
    // cost of the winner without penalties: state part of its rollout + R-only control polynomial
    d.cost_adm[c.b] = scs[bi][c.lane] + fma(al, fma(al, d.cq[5 * S + c.b], d.cq[4 * S + c.b]), d.cq[3 * S + c.b]);
    admm_finish<M>(d, c, outer, inner, bi, sqrt(ps), sqrt(ds));
  }
  return nchunks;
}

template <class M, int CPT, int MAXW, int MINB, bool PX>
__global__ void __launch_bounds__(TILE * MAXW, MINB) k_linesearch(Dev d, int fuse, int outer, int inner) {
  ls_tile<M, CPT, MAXW, PX>(d, d.tile0 + blockIdx.x * d.tstep, fuse, outer, inner, 0u, true);
}

// Persistent form for the overlapped schedule (see ModelImpl::ilqr_admm): exactly as many CTAs as stay resident next
// to the HBM-bound feed-forward kernel of the other half batch; tiles are fetched from a counter (ctr[0]; the last CTA
// to leave resets it and the exit count ctr[1], so the pair is ready for the next launch).
template <class M, int CPT, int MAXW, int MINB, bool PX>
__global__ void __launch_bounds__(TILE * MAXW, MINB) k_linesearch_pers(Dev d, int fuse, int outer, int inner, int *ctr) {
  __shared__ int s_tile;
  const bool t0 = threadIdx.x == 0 && threadIdx.y == 0;
  const int ntiles = (d.tile1 - d.tile0 + d.tstep - 1) / d.tstep;
  unsigned gch = 0;                                           // 0 until a tile got past its skip test = barriers not initialised
  for (;;) {
    __syncthreads();                                          // the previous tile's shared state is dead
    if (t0) s_tile = atomicAdd(&ctr[0], 1);
    __syncthreads();
    const int i = s_tile;
    if (i >= ntiles) break;
    gch += (unsigned)ls_tile<M, CPT, MAXW, PX>(d, d.tile0 + i * d.tstep, fuse, outer, inner, gch, gch == 0);
  }
  if (t0 && atomicAdd(&ctr[1], 1) == (int)gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; }
}

// ---- The whole inner ADMM loop in ONE launch (control-only projections, large batches).  A persistent CTA owns the
// tiles blockIdx.x, blockIdx.x + gridDim.x, ... for all ADMM iterations of an outer iteration (nothing couples tiles
// inside one: admm.py:31-97 runs per problem).  Per ADMM iteration and round of up to W tiles: every warp runs the
// TMA-staged ff-pass + linear rollout of ONE tile (ff_tma_tile, private operand ring: HBM-bound, W x 3 streaming warps
// per SM like the stand-alone k_ff_tma), then the whole CTA runs the line search + streaming ADMM update of these tiles
// one after the other (ls_tile: FP64-bound).  The CTAs of an SM drift apart in phase, so the HBM-bound half of one
// overlaps the FP64-bound half of its neighbours - which two separate kernels never did: the line search needs the
// whole register file for its residency (profiles/r2_tuning_log.md).  Same device functions, same arithmetic, same order
// per problem as the k_ff_tma / k_linesearch launch pair: results are bit-identical.  `stagger`: the s-th CTA to arrive on
// an SM (per-SM counter) sleeps s x stagger ns once, so the first wave does not start in lock-step.
template <class M, int CPT, int MAXW, int MINB, bool JC, int TC, int NST>
__global__ void __launch_bounds__(TILE * MAXW, MINB) k_admm_loop(Dev d, int outer, int a0, int a1, unsigned stagger,
                                                                  int *sm_ctr) {
  using FS = FfTmaShape<M, false, JC, TC, NST>;
  extern __shared__ __align__(128) double smem_fft[];
  const int w = threadIdx.y, lane = threadIdx.x, W = blockDim.y;
  constexpr size_t RSTRIDE = (128 + FS::RING_BYTES) / sizeof(double);     // per-warp mbarriers + ring
  double *my = smem_fft + (size_t)w * RSTRIDE;
  double *cst = smem_fft + (size_t)W * RSTRIDE;                            // one copy of the plan constants
  ff_tma_setup<M, false, JC, TC, NST>(d, my, lane, cst, w == 0);
  if (stagger && w == 0) {
    unsigned slot = 0;
    if (lane == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      slot = (unsigned)atomicAdd(&sm_ctr[smid & 255], 1);
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < 3 && slot > 0) {                       // first wave only
      for (unsigned left = stagger * slot; left > 0;) {
        const unsigned q = left > 100000u ? 100000u : left;
        __nanosleep(q);
        left -= q;
      }
    }
  }
  __syncthreads();
  const int ntiles = (d.tile1 - d.tile0 + d.tstep - 1) / d.tstep, G = gridDim.x;
  const int mine = (ntiles - (int)blockIdx.x + G - 1) / G;              // tiles of this CTA
  const int rounds = (mine + W - 1) / W;
  auto tile_of = [&](int k) { return d.tile0 + ((int)blockIdx.x + k * G) * d.tstep; };
  unsigned gb = 0, gf = 0, gch = 0;
  for (int a = a0; a < a1; a++) {
    bool any = false;
    int k0 = 0;
    for (int r = 0; r < rounds; r++) {
      const int cnt = mine / rounds + (r < mine % rounds ? 1 : 0);      // rounds of (nearly) equal size
      if (w < cnt) {
        ff_tma_tile<M, false, JC, TC, NST>(d, my, tile_of(k0 + w), lane, gb, gf, cst);
        fence_proxy_async();      // du: generic-proxy stores of this warp, read by the line search's bulk copies
      }
      __syncthreads();
      for (int q = 0; q < cnt; q++) {
        const int used = ls_tile<M, CPT, MAXW, false, 5>(d, tile_of(k0 + q), 1, outer, a, gch, gch == 0);
        gch += (unsigned)used;
        any |= used != 0;
        fence_proxy_async();      // reg_u: generic-proxy stores of the epilogue, read by the next ff-pass's bulk copies
        __syncthreads();          // (also: the tile's shared state is dead)
      }
      k0 += cnt;
    }
    if (!any) break;              // every problem of this CTA has left the ADMM loop (uniform over the CTA)
  }
}

// ---- state projection onto the outside of obstacle sets: project_set_convex (isls/projections.py:289-374) with
// As = I, bs = 0 and per-set projections  p -> c + W^-1 Pi_sq(W (p - c))  on the position (project_square_batch,
// projections.py:246-255; Car/Iterative LQR with state constraints.ipynb cell 18).  One problem per thread; all N rows
// are projected together (the reference's stop rule is the maximum over sets and rows): k_admm parks the winner's
// states and the pre-projection points, k_obst_project (one CTA per problem, one thread per row, z_k / lambda_k in
// registers, block reductions for the stop rule) projects and finishes the ADMM update.
__device__ __forceinline__ void obst_project_one(const Dev &d, int k, double (&y)[2]) {
  const double c0 = d.ob_c[2 * k], c1 = d.ob_c[2 * k + 1];
  if (d.obst_kind == 1) {        // project_quadratic_batch(p - c, lower, upper) + c   (projections.py:91-105)
    const double z0 = y[0] - c0, z1 = y[1] - c1;
    const double ss = z0 * z0 + z1 * z1, val = 0.5 * ss;
    double s2 = 0.0;                                                     // x * sqrt(2 l) / ||x||, left to right
    if (val > d.obst_upper) s2 = sqrt(2.0 * d.obst_upper);
    if (d.ob_lo[k] > val) s2 = sqrt(2.0 * d.ob_lo[k]);                   // later mask wins, like the numpy code
    const double nr = sqrt(ss);
    y[0] = (s2 == 0.0 ? z0 : (z0 * s2) / nr) + c0;
    y[1] = (s2 == 0.0 ? z1 : (z1 * s2) / nr) + c1;
    return;
  }
  const double *W = d.ob_W + 4 * k, *Wi = d.ob_Wi + 4 * k;
  const double z0 = y[0] - c0, z1 = y[1] - c1;
  double w0 = z0 * W[0] + z1 * W[1], w1 = z0 * W[2] + z1 * W[3];          // z @ W.T
  const double lo = d.ob_lo[k], up = d.obst_upper;
  const double a0 = fabs(w0), a1 = fabs(w1);
  if (fmax(a0, a1) < lo) {                                                // np.argmax: first maximum
    if (a0 >= a1) w0 = lo * ((w0 > 0.0) - (w0 < 0.0));
    else w1 = lo * ((w1 > 0.0) - (w1 < 0.0));
  }
  w0 = fmax(fmin(w0, up), -up);
  w1 = fmax(fmin(w1, up), -up);
  y[0] = (w0 * Wi[0] + w1 * Wi[1]) + c0;                                  // zp @ W_inv.T + c
  y[1] = (w0 * Wi[2] + w1 * Wi[3]) + c1;
}

// Winner rollout + ADMM update: re-roll the chosen candidate (the primal iterate (x,u) returned by f_argmin,
// isls/isls.py:478; it is not stored - k_outer_end re-rolls the last one in place), apply the z-projection and scaled dual update element by element (admm.py:43-59), form the
// residual norms (admm.py:62-69) and run the stop tests (admm.py:72-85).
// Operand fetch policies of admm_body.  AdmmFetchGlobal: plain global loads, all of a step issued together (one memory
// round trip per step).  AdmmFetchStaged (k_admm_staged, small batches): each thread streams its own operands of the
// next STAGES-1 steps into a private slice of shared memory with cp.async, like k_ff_staged.
template <class M>
struct AdmmFetchGlobal {
  static constexpr int n = M::n, m = M::m;
  const Dev &d;
  const double *uh, *du, *zx, *lx, *zu, *lu;
  __device__ __forceinline__ AdmmFetchGlobal(const Dev &d_, const TileCtx<M> &c)
      : d(d_), uh(c.at(d_.uh, d_, m)), du(c.at(d_.du, d_, m)), zx(c.at(d_.zx, d_, n)), lx(c.at(d_.lx, d_, n)),
        zu(c.at(d_.zu, d_, m)), lu(c.at(d_.lu, d_, m)) {}
  __device__ __forceinline__ void operator()(int t, double (&duv)[m], double (&uhv)[m], double (&zuv)[m],
                                             double (&luv)[m], double (&zxv)[n], double (&lxv)[n]) {
#pragma unroll
    for (int j = 0; j < m; j++) {
      duv[j] = EL(du, m, t, j);
      uhv[j] = EL(uh, m, t, j);
      if (d.proj_u) { zuv[j] = EL(zu, m, t, j); luv[j] = EL(lu, m, t, j); }
    }
    if (d.proj_x) {
#pragma unroll
      for (int i = 0; i < n; i++) { zxv[i] = EL(zx, n, t, i); lxv[i] = EL(lx, n, t, i); }
    }
  }
};

template <class M, int STAGES>
struct AdmmFetchStaged {
  static constexpr int n = M::n, m = M::m, SL = 4 * m + 2 * n;   // du, u^, z_u, lambda_u, z_x, lambda_x
  const Dev &d;
  const double *uh, *du, *zx, *lx, *zu, *lu;
  double *sm;
  int t_issue;
  __device__ __forceinline__ double *slot(int stage, int k) const { return sm + ((size_t)stage * SL + k) * TILE; }
  __device__ __forceinline__ void issue(int t) {
    const int st = t % STAGES;
    int k = 0;
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slot(st, k++), &EL(du, m, t, j));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slot(st, k++), &EL(uh, m, t, j));
    if (d.proj_u) {
#pragma unroll
      for (int j = 0; j < m; j++) { cp_async8(slot(st, k + j), &EL(zu, m, t, j)); cp_async8(slot(st, k + m + j), &EL(lu, m, t, j)); }
    }
    k += 2 * m;
    if (d.proj_x) {
#pragma unroll
      for (int i = 0; i < n; i++) { cp_async8(slot(st, k + i), &EL(zx, n, t, i)); cp_async8(slot(st, k + n + i), &EL(lx, n, t, i)); }
    }
  }
  __device__ __forceinline__ AdmmFetchStaged(const Dev &d_, const TileCtx<M> &c, double *smem)
      : d(d_), uh(c.at(d_.uh, d_, m)), du(c.at(d_.du, d_, m)), zx(c.at(d_.zx, d_, n)), lx(c.at(d_.lx, d_, n)),
        zu(c.at(d_.zu, d_, m)), lu(c.at(d_.lu, d_, m)), sm(smem + c.lane), t_issue(0) {
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (t_issue < d.N) issue(t_issue);
      cp_async_commit();
      t_issue++;
    }
  }
  __device__ __forceinline__ void operator()(int t, double (&duv)[m], double (&uhv)[m], double (&zuv)[m],
                                             double (&luv)[m], double (&zxv)[n], double (&lxv)[n]) {
    if (t_issue < d.N) issue(t_issue);
    cp_async_commit();
    t_issue++;
    cp_async_wait<STAGES - 1>();
    const int st = t % STAGES;
    int k = 0;
#pragma unroll
    for (int j = 0; j < m; j++) duv[j] = *slot(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) uhv[j] = *slot(st, k++);
    if (d.proj_u) {
#pragma unroll
      for (int j = 0; j < m; j++) { zuv[j] = *slot(st, k + j); luv[j] = *slot(st, k + m + j); }
    }
    k += 2 * m;
    if (d.proj_x) {
#pragma unroll
      for (int i = 0; i < n; i++) { zxv[i] = *slot(st, k + i); lxv[i] = *slot(st, k + n + i); }
    }
  }
};

template <class M, class Fetch>
__device__ __forceinline__ void admm_body(const Dev &d, const TileCtx<M> &c, int outer, int inner, Fetch &fetch) {
  constexpr int n = M::n, m = M::m;
  const int tile = c.tile;
  const double *xh = c.at(d.xh, d, n);
  double *zx = c.at(d.zx, d, n), *lx = c.at(d.lx, d, n), *rgx = c.at(d.rgx, d, n);
  double *zu = c.at(d.zu, d, m), *lu = c.at(d.lu, d, m), *rgu = c.at(d.rgu, d, m);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  const int bi = d.best[c.b];
  const double al = d.alphas[bi];
  double x[n], u[m], xn[n];
#pragma unroll
  for (int i = 0; i < n; i++) x[i] = EL(xh, n, 0, i);
  double cs = 0.0, cc = 0.0, prx = 0.0, pru = 0.0, drx = 0.0, dru = 0.0;
  int8_t *mkx = (d.out.mask_x && c.valid) ? d.out.mask_x + c.ob * d.N * n : nullptr;
  int8_t *mku = (d.out.mask_u && c.valid) ? d.out.mask_u + c.ob * d.N * m : nullptr;
  // bounds of a step: read-only plan constants, fetched one step ahead (re-loaded in place after their last use)
  double lou[m], hiu[m], lox[n], hix[n];
  int qnz_t = 0;
  auto ld_bounds = [&](int t) {
    qnz_t = __ldg(d.qnz + t);
    if (d.proj_u) {
#pragma unroll
      for (int j = 0; j < m; j++) { lou[j] = __ldg(d.lo_u + t * m + j); hiu[j] = __ldg(d.hi_u + t * m + j); }
    }
    if (d.proj_x && d.n_obst == 0) {
#pragma unroll
      for (int i = 0; i < n; i++) { lox[i] = __ldg(d.lo_x + t * n + i); hix[i] = __ldg(d.hi_x + t * n + i); }
    }
  };
  ld_bounds(0);
  for (int t = 0; t < d.N; t++) {
    // all loads of the step are issued before the first store (the stores may alias as far as the compiler knows,
    // which would otherwise serialise one memory round trip per element)
    // (control-side AND state-side: with the control u formed between them the state-side loads were only issued
    // after the control-side ones had arrived - two serialised round trips per step, 54 % of k_admm<Arm3Model> in the
    // ncu source view, profiles/r1_c3_small_batch_kernels.md)
    double duv[m], uhv[m], zuv[m], luv[m], zxv[n], lxv[n];
    fetch(t, duv, uhv, zuv, luv, zxv, lxv);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = fma(al, duv[j], uhv[j]);
    cc += ctrl_sq<M>(d, u);
#pragma unroll
    for (int j = 0; j < m; j++) {
      if (d.proj_u) {
        int mk;
        admm_elem(u[j], d.relax, lou[j], hiu[j], zuv[j], luv[j], pru, dru, mk);
        EL(zu, m, t, j) = zuv[j];
        EL(lu, m, t, j) = luv[j];
        EL(rgu, m, t, j) = __dsub_rn(zuv[j], luv[j]);   // reg = z - lambda for the next f_argmin (admm.py:32-33)
        if (mku) mku[t * m + j] = (int8_t)mk;
      }
    }
    if (d.proj_x && d.n_obst > 0) {
      // obstacle sets: park the winner state and the pre-projection point; the rows are projected together below
      const size_t arr = (size_t)d.T * d.N * n * TILE;
      double *xw = c.at(d.obw, d, n), *pre = c.at(d.obw + arr, d, n);
#pragma unroll
      for (int i = 0; i < n; i++) {
        EL(xw, n, t, i) = x[i];
        EL(pre, n, t, i) = __dadd_rn(__dadd_rn(__dmul_rn(d.relax, x[i]), __dmul_rn(__dsub_rn(1.0, d.relax), zxv[i])), lxv[i]);
      }
    } else if (d.proj_x) {
#pragma unroll
      for (int i = 0; i < n; i++) {
        int mk;
        admm_elem(x[i], d.relax, lox[i], hix[i], zxv[i], lxv[i], prx, drx, mk);
        EL(zx, n, t, i) = zxv[i];
        EL(lx, n, t, i) = lxv[i];
        EL(rgx, n, t, i) = __dsub_rn(zxv[i], lxv[i]);
        if (mkx) mkx[t * n + i] = (int8_t)mk;
      }
    }
    cs += state_cost<M>(d, zs, t, x, qnz_t);
    if (t + 1 < d.N) ld_bounds(t + 1);
    M::step(x, u, xn, d.dt);
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = xn[i];
  }
  d.cost_adm[c.b] = cs + d.u_std * cc;
  if (d.proj_x && d.n_obst > 0) {                         // k_obst_project finishes: park the control-side residual
    const size_t S = (size_t)d.T * TILE;                  // sums in two cq rows (free between the line search and k_ff)
    d.cq[3 * S + c.b] = pru;
    d.cq[4 * S + c.b] = dru;
    return;
  }
  admm_finish<M>(d, c, outer, inner, bi, sqrt(prx) + sqrt(pru), sqrt(drx) + sqrt(dru));   // admm.py:62-69
}

// residual log + ADMM stop tests (admm.py:62-97) of one problem after its z / lambda update
template <class M>
__device__ __forceinline__ void admm_finish(const Dev &d, const TileCtx<M> &c, int outer, int inner, int bi,
                                            double prim, double dual) {
  const double pprim = d.prim[c.b], pdual = d.dual[c.b];
  d.prim[c.b] = prim;
  d.dual[c.b] = dual;
  d.ait[c.b] = inner + 1;
  if (c.valid) {
    if (d.out.res_log) {
      double *r = d.out.res_log + (((size_t)c.ob * d.max_outer + outer) * d.max_admm + inner) * 2;
      r[0] = prim;
      r[1] = dual;
    }
    if (d.out.alpha_idx) d.out.alpha_idx[((size_t)c.ob * d.max_outer + outer) * d.max_admm + inner] = bi;
    if (d.out.admm_iters) d.out.admm_iters[c.ob * d.max_outer + outer] = inner + 1;
  }
  int ex = 0;
  if (!d.fixed_budget) {
    if (prim < d.tol && dual < d.tol) ex = ISLS_ADMM_CONVERGED;                  // admm.py:72
    else {
      const double pch = fabs(pprim - prim) / (pprim + 1e-30);                   // admm.py:78-79
      const double dch = fabs(pdual - dual) / (pdual + 1e-30);
      if (pch < d.stall_tol && dch < d.stall_tol) ex = ISLS_ADMM_STALLED;        // admm.py:80
    }
  }
  if (!ex && inner == d.max_admm - 1) ex = ISLS_ADMM_MAXIT;
  if (ex) {
    d.adone[c.b] = 1;
    if (c.valid && d.out.admm_exit) d.out.admm_exit[c.ob * d.max_outer + outer] = ex;
  }
}

template <class M>
__global__ void k_admm(Dev d, int outer, int inner) {
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b] || d.adone[c.b]) return;
  AdmmFetchGlobal<M> fetch(d, c);
  admm_body<M>(d, c, outer, inner, fetch);
}

// small batches (latency-bound, < 1,536 tiles): one tile per CTA, per-step operands cp.async-staged STAGES deep
template <class M, int STAGES>
__global__ void __launch_bounds__(TILE) k_admm_staged(Dev d, int outer, int inner) {
  extern __shared__ double smem_admm[];
  const int tile = d.tile0 + blockIdx.x;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b] || d.adone[c.b]) return;
  AdmmFetchStaged<M, STAGES> fetch(d, c, smem_admm);
  admm_body<M>(d, c, outer, inner, fetch);
}

template <class M>
static int launch_admm(const Dev &d, int outer, int inner, cudaStream_t s) {
  static int mode = -2;
  if (mode == -2) {
    const char *e = getenv("ISLS_ADMM_STAGES");       // -1 auto (default), 0 plain, 6 forced
    mode = e ? atoi(e) : -1;
  }
  const int tiles = d.tile1 - d.tile0;
  constexpr int ST = 6;     // the step is short: at 4 stages 7 % of the samples still sat on the cp.async wait
  const size_t smem = (size_t)ST * AdmmFetchStaged<M, ST>::SL * TILE * sizeof(double);
  const bool staged = mode == ST || (mode < 0 && tiles < 1536 && smem <= 48 * 1024);
  if (staged) {
    k_admm_staged<M, ST><<<tiles, TILE, smem, s>>>(d, outer, inner);
  } else {
    k_admm<M><<<dim3((tiles + 1) / 2), dim3(TILE, 2), 0, s>>>(d, outer, inner);
  }
  return 0;
}

// ------------------------------------------------------------------------------ robust iSLS-ADMM (isls.py:503-712)
// Riccati form of  [d_u | Phi_u] = l_side^-1 (r_side + Rr reg)  (isls.py:562-579), column by column: column 0 is the
// iLQR-ADMM step (k_ff: cx = 2Q(x^ - z), cu = 2R u^ + 2Rr(u^ - reg_abs) with reg_abs = u^ + reg_0, dx_0 = 0); column
// c >= 1 has dx_0 = e_c (Sx = C[:, :dim]), cx = 0, cu = -2Rr reg_c.

// start of an outer iteration: lambda = 0 (isls.py:615), z warm start kept in delta coordinates (isls.py:695-696),
// reg_abs of column 0 for k_ff
template <class M>
__global__ void k_isls_reset(Dev d) {
  constexpr int m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b]) return;
  const int C = d.isls_C;
  const double *uh = c.at(d.uh, d, m), *Zm = c.at(d.Zm, d, m * C);
  double *Lm = c.at(d.Lm, d, m * C), *rgu = c.at(d.rgu, d, m);
  for (int t = 0; t < d.N; t++)
    for (int j = 0; j < m; j++) {
      for (int q = 0; q < C; q++) EL(Lm, m * C, t, j * C + q) = 0.0;
      EL(rgu, m, t, j) = EL(uh, m, t, j) + EL(Zm, m * C, t, j * C);
    }
  if (d.Zx) {                                            // lmb_x = 0 (isls.py:614), reg_abs of column 0
    constexpr int n = M::n;
    const double *xh = c.at(d.xh, d, n), *Zx = c.at(d.Zx, d, n * C);
    double *Lx = c.at(d.Lx, d, n * C), *rgx = c.at(d.rgx, d, n);
    for (int t = 0; t < d.N; t++)
      for (int i = 0; i < n; i++) {
        for (int q = 0; q < C; q++) EL(Lx, n * C, t, i * C + q) = 0.0;
        EL(rgx, n, t, i) = EL(xh, n, t, i) + EL(Zx, n * C, t, i * C);
      }
  }
}

// columns c = 1..dim (blockIdx.y + 1): feed-forward sweep (sls.py:168-202) with cx = 0, cu = -2 Rr (z_c - lambda_c),
// batch-form last control, then the linear rollout from dx_0 = e_c.  k_t is parked in Xu[., c] between the sweeps.
template <class M>
__global__ void k_isls_cols(Dev d) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m);
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b] || d.adone[c.b]) return;
  const int C = d.isls_C, col = blockIdx.y + 1, N = d.N;
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  if (col == C) {
    // extra job of the state-side variant: x_x[:, 0] = x_noms[ind] - x_nom (isls.py:605-606), the winner's open-loop
    // rollout (same arithmetic as k_outer_end, which makes it the next nominal trajectory)
    const double *du = c.at(d.du, d, m);
    double *Xx = c.at(d.Xx, d, n * C);
    const double al = d.alphas[d.best[c.b]];
    double x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, 0, i);
    for (int t = 0; t < N; t++) {
#pragma unroll
      for (int j = 0; j < m; j++) u[j] = fma(al, EL(du, m, t, j), EL(uh, m, t, j));
#pragma unroll
      for (int i = 0; i < n; i++) EL(Xx, n * C, t, i * C) = x[i] - EL(xh, n, t, i);
      M::step(x, u, xn, d.dt);
#pragma unroll
      for (int i = 0; i < n; i++) x[i] = xn[i];
    }
    return;
  }
  const double *Kg = c.at(d.Kg, d, m * n), *Qx = c.at(d.Qux, d, m * n);
  const double *Qu = c.at(d.Quu, d, nt), *Qi = c.at(d.Qui, d, nt);
  const double *Zm = c.at(d.Zm, d, m * C), *Lm = c.at(d.Lm, d, m * C);
  double *Xu = c.at(d.Xu, d, m * C);
  const double *Zx = d.Zx ? c.at(d.Zx, d, n * C) : nullptr, *Lx = d.Zx ? c.at(d.Lx, d, n * C) : nullptr;
  double *Xx = d.Zx ? c.at(d.Xx, d, n * C) : nullptr;
  double A[n][n], Bm[n][m], v[n], cx[n];
  init_AB<M>(A, Bm);
  // state side (isls.py:571-572 in Riccati form): the column's linear state term cx = -2 Qr (z_x,c - lambda_x,c)
  auto cx_of = [&](int t, double (&cxv)[n]) {
#pragma unroll
    for (int i = 0; i < n; i++)
      cxv[i] = Zx ? -2.0 * __ldg(d.rho_x + t * n + i) * (EL(Zx, n * C, t, i * C + col) - EL(Lx, n * C, t, i * C + col))
                  : 0.0;
  };
  cx_of(N - 1, cx);
#pragma unroll
  for (int i = 0; i < n; i++) v[i] = cx[i];
  auto cu_of = [&](int t, double (&cu)[m]) {
#pragma unroll
    for (int j = 0; j < m; j++)
      cu[j] = -2.0 * __ldg(d.rho_u + t * m + j) * (EL(Zm, m * C, t, j * C + col) - EL(Lm, m * C, t, j * C + col));
  };
  {
    double cu[m];
    cu_of(N - 1, cu);
#pragma unroll
    for (int j = 0; j < m; j++)
      EL(Xu, m * C, N - 1, j * C + col) = -cu[j] / (2.0 * (d.u_std * d.Rw[j] + d.rho_u[(N - 1) * m + j]));
  }
  for (int t = N - 2; t >= 0; t--) {
    double x[n], u[m], J[M::NJA], cu[m], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, t, i);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = EL(uh, m, t, j);
#pragma unroll
    for (int a = 0; a < m; a++) {
#pragma unroll
      for (int j = 0; j < n; j++) Qux[a][j] = EL(Qx, m * n, t, a * n + j);
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) {
        Qui[a][b2] = EL(Qi, nt, t, tri(a, b2)); Qui[b2][a] = Qui[a][b2];
        Quu[a][b2] = EL(Qu, nt, t, tri(a, b2)); Quu[b2][a] = Quu[a][b2];
      }
    }
    cu_of(t, cu);                  // its loads go out with the step's other loads, not after the Jacobian has waited
    cx_of(t, cx);
    M::jac(x, u, J, d.dt);
    M::expand(J, A, Bm, d.dt);
    ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
#pragma unroll
    for (int j = 0; j < m; j++) EL(Xu, m * C, t, j * C + col) = kt[j];
  }
  double dx[n];
#pragma unroll
  for (int i = 0; i < n; i++) dx[i] = (i == col - 1) ? 1.0 : 0.0;
  for (int t = 0; t < N; t++) {
    double x[n], u[m], K[m][n], kv[m], duv[m];
    if (Xx) {                      // x_x[:, c] = Sx[:, c] + Su du_c (isls.py:580-582)
#pragma unroll
      for (int i = 0; i < n; i++) EL(Xx, n * C, t, i * C + col) = dx[i];
    }
#pragma unroll
    for (int a = 0; a < m; a++) {
#pragma unroll
      for (int j = 0; j < n; j++) K[a][j] = EL(Kg, m * n, t, a * n + j);
      kv[a] = EL(Xu, m * C, t, a * C + col);
      u[a] = EL(uh, m, t, a);
    }
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, t, i);
#pragma unroll
    for (int a = 0; a < m; a++) {
      double acc = 0.0;
      if (t < N - 1) {
#pragma unroll
        for (int j = 0; j < n; j++) acc = fma(K[a][j], dx[j], acc);
      }
      duv[a] = acc + kv[a];
      EL(Xu, m * C, t, a * C + col) = duv[a];
    }
    if (t < N - 1) {
      double J[M::NJA], dxn[n];
      M::jac(x, u, J, d.dt);
      M::expand(J, A, Bm, d.dt);
      mat_Ax_Bu<M>(A, Bm, dx, duv, dxn);
#pragma unroll
      for (int i = 0; i < n; i++) dx[i] = dxn[i];
    }
  }
}

// cp.async-staged form of k_isls_cols for small batches (one warp per CTA, each thread streams its own operands of the
// next STAGES-1 steps into a private slice of shared memory, as in k_ff_staged); identical arithmetic.
template <class M, int STAGES>
__global__ void __launch_bounds__(TILE) k_isls_cols_staged(Dev d) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m);
  constexpr int SB = n + m + m * n + 2 * nt + 2 * m;     // backward slots: x^, u^, Qux, Quu, Quu^-1, Z_c, Lambda_c
  constexpr int SF = m * n + 2 * m + n;                  // forward slots:  K, k_c, u^, x^
  constexpr int SFW = (STAGES * SB) / SF < 8 ? (STAGES * SB) / SF : 8;
  extern __shared__ double smem_cols[];
  const int tile = d.tile0 + blockIdx.x;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b] || d.adone[c.b]) return;
  const int tid = threadIdx.x;
  const int C = d.isls_C, col = blockIdx.y + 1, N = d.N;
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *Kg = c.at(d.Kg, d, m * n), *Qx = c.at(d.Qux, d, m * n);
  const double *Qu = c.at(d.Quu, d, nt), *Qi = c.at(d.Qui, d, nt);
  const double *Zm = c.at(d.Zm, d, m * C), *Lm = c.at(d.Lm, d, m * C);
  double *Xu = c.at(d.Xu, d, m * C);
  auto slot = [&](int stage, int k) -> double * { return smem_cols + ((size_t)stage * SB + k) * TILE + tid; };
  auto slotf = [&](int stage, int k) -> double * { return smem_cols + ((size_t)stage * SF + k) * TILE + tid; };
  double A[n][n], Bm[n][m], v[n], cx[n];
  init_AB<M>(A, Bm);
#pragma unroll
  for (int i = 0; i < n; i++) { v[i] = 0.0; cx[i] = 0.0; }
  {
#pragma unroll
    for (int j = 0; j < m; j++) {
      const double rho = d.rho_u[(N - 1) * m + j];
      const double cu = -2.0 * rho * (EL(Zm, m * C, N - 1, j * C + col) - EL(Lm, m * C, N - 1, j * C + col));
      EL(Xu, m * C, N - 1, j * C + col) = -cu / (2.0 * (d.u_std * d.Rw[j] + rho));
    }
  }
  auto issue_b = [&](int t, int stage) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < n; i++) cp_async8(slot(stage, k++), &EL(xh, n, t, i));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slot(stage, k++), &EL(uh, m, t, j));
#pragma unroll
    for (int q = 0; q < m * n; q++) cp_async8(slot(stage, k++), &EL(Qx, m * n, t, q));
#pragma unroll
    for (int q = 0; q < nt; q++) cp_async8(slot(stage, k++), &EL(Qu, nt, t, q));
#pragma unroll
    for (int q = 0; q < nt; q++) cp_async8(slot(stage, k++), &EL(Qi, nt, t, q));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slot(stage, k++), &EL(Zm, m * C, t, j * C + col));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slot(stage, k++), &EL(Lm, m * C, t, j * C + col));
  };
  int t_issue = N - 2;
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (t_issue >= 0) issue_b(t_issue, (N - 2 - t_issue) % STAGES);
    cp_async_commit();
    t_issue--;
  }
  double rho_c[m];
#pragma unroll
  for (int j = 0; j < m; j++) rho_c[j] = N >= 2 ? __ldg(d.rho_u + (N - 2) * m + j) : 0.0;
  for (int t = N - 2; t >= 0; t--) {
    if (t_issue >= 0) issue_b(t_issue, (N - 2 - t_issue) % STAGES);
    cp_async_commit();
    t_issue--;
    cp_async_wait<STAGES - 1>();
    const int st = (N - 2 - t) % STAGES;
    double x[n], u[m], J[M::NJA], cu[m], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
    int k = 0;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = *slot(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = *slot(st, k++);
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int j = 0; j < n; j++) Qux[a][j] = *slot(st, k++);
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) { Quu[a][b2] = *slot(st, k + tri(a, b2)); Quu[b2][a] = Quu[a][b2]; }
    k += nt;
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int b2 = 0; b2 <= a; b2++) { Qui[a][b2] = *slot(st, k + tri(a, b2)); Qui[b2][a] = Qui[a][b2]; }
    k += nt;
#pragma unroll
    for (int j = 0; j < m; j++) cu[j] = -2.0 * rho_c[j] * (*slot(st, k + j) - *slot(st, k + m + j));
#pragma unroll
    for (int j = 0; j < m; j++) rho_c[j] = t > 0 ? __ldg(d.rho_u + (t - 1) * m + j) : 0.0;
    M::jac(x, u, J, d.dt);
    M::expand(J, A, Bm, d.dt);
    ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
#pragma unroll
    for (int j = 0; j < m; j++) EL(Xu, m * C, t, j * C + col) = kt[j];
  }
  cp_async_wait<0>();
  auto issue_f = [&](int t, int stage) {
    int k = 0;
#pragma unroll
    for (int q = 0; q < m * n; q++) cp_async8(slotf(stage, k++), &EL(Kg, m * n, t, q));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slotf(stage, k++), &EL(Xu, m * C, t, j * C + col));
#pragma unroll
    for (int j = 0; j < m; j++) cp_async8(slotf(stage, k++), &EL(uh, m, t, j));
#pragma unroll
    for (int i = 0; i < n; i++) cp_async8(slotf(stage, k++), &EL(xh, n, t, i));
  };
  double dx[n];
#pragma unroll
  for (int i = 0; i < n; i++) dx[i] = (i == col - 1) ? 1.0 : 0.0;
  t_issue = 0;
#pragma unroll
  for (int s = 0; s < SFW - 1; s++) {
    if (t_issue < N) issue_f(t_issue, t_issue % SFW);
    cp_async_commit();
    t_issue++;
  }
  for (int t = 0; t < N; t++) {
    if (t_issue < N) issue_f(t_issue, t_issue % SFW);
    cp_async_commit();
    t_issue++;
    cp_async_wait<SFW - 1>();
    const int st = t % SFW;
    double x[n], u[m], K[m][n], kv[m], duv[m];
    int k = 0;
#pragma unroll
    for (int a = 0; a < m; a++)
#pragma unroll
      for (int j = 0; j < n; j++) K[a][j] = *slotf(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) kv[j] = *slotf(st, k++);
#pragma unroll
    for (int j = 0; j < m; j++) u[j] = *slotf(st, k++);
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = *slotf(st, k++);
#pragma unroll
    for (int a = 0; a < m; a++) {
      double acc = 0.0;
      if (t < N - 1) {
#pragma unroll
        for (int j = 0; j < n; j++) acc = fma(K[a][j], dx[j], acc);
      }
      duv[a] = acc + kv[a];
      EL(Xu, m * C, t, a * C + col) = duv[a];
    }
    if (t < N - 1) {
      double J[M::NJA], dxn[n];
      M::jac(x, u, J, d.dt);
      M::expand(J, A, Bm, d.dt);
      mat_Ax_Bu<M>(A, Bm, dx, duv, dxn);
#pragma unroll
      for (int i = 0; i < n; i++) dx[i] = dxn[i];
    }
  }
  cp_async_wait<0>();
}

template <class M>
static void launch_isls_cols(const Dev &d, cudaStream_t s) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m), SB = n + m + m * n + 2 * nt + 2 * m, ST = 3;
  const int tiles = d.tile1 - d.tile0, cols = d.isls_C - 1;
  const size_t smem = (size_t)ST * SB * TILE * sizeof(double);
  static int mode = -2;
  if (mode == -2) {
    const char *e = getenv("ISLS_COLS_STAGES");       // -1 auto (default), 0 plain, 3 forced
    mode = e ? atoi(e) : -1;
  }
  if (!d.Zx && (mode == ST || (mode < 0 && (long long)tiles * cols < 1536 && smem <= 48 * 1024))) {
    k_isls_cols_staged<M, ST><<<dim3(tiles, cols), TILE, smem, s>>>(d);
  } else {                          // with a state side: one more job per tile (the winner's rollout, column 0 of x_x)
    k_isls_cols<M><<<dim3((tiles + 1) / 2, cols + (d.Zx ? 1 : 0)), dim3(TILE, 2), 0, s>>>(d);
  }
}

// ADMM update on the matrix variable (isls.py:628-654): one CTA per problem, thread r = row (t, j) of
// [d_u | Phi_u(:, :dim)]; z = project_u(alpha x + (1 - alpha) z + lambda, u_nom) with the notebook's closure (column 0
// shifted by u_nom, project_set_convex over the SOC set, shifted back), lambda += x - z, residuals weighted by Rr.
// Launch bounds: the cone-shape-specialised forms keep z_i, lambda_i in registers (<= 384 threads = N m <= 384 rows);
// the run-time form serves any shape and up to 1,024 rows.
template <class M, int CP = 0, int CC = 0, int CR = 0>
__global__ void __launch_bounds__(CP ? 384 : 1024) k_isls_update(Dev d, SocSet S, SocX X, int outer, int inner) {
  constexpr int m = M::m;
  __shared__ double red[32];
  const long long b = blockIdx.x;
  const int tile = (int)(b / TILE), lane = (int)(b % TILE);
  TileCtx<M> c(d, tile, lane);
  if (!c.valid || d.odone[c.b] || d.adone[c.b]) return;               // uniform over the CTA
  const int C = CC ? CC : d.isls_C, r = threadIdx.x, rows = d.N * m;
  const bool act = r < rows;
  const int t = act ? r / m : 0, j = act ? r % m : 0;
  const double *uh = c.at(d.uh, d, m), *du = c.at(d.du, d, m);
  double *Zm = c.at(d.Zm, d, m * C), *Lm = c.at(d.Lm, d, m * C), *Xu = c.at(d.Xu, d, m * C), *rgu = c.at(d.rgu, d, m);
  const int bi = d.best[c.b];
  double xu[SOC_MAXC] = {}, z[SOC_MAXC] = {}, lm[SOC_MAXC] = {}, y[SOC_MAXC] = {}, zn[SOC_MAXC] = {};
  double un = 0.0, rho = 0.0;
  if (act) {
    un = EL(uh, m, t, j);
    rho = d.rho_u[t * m + j];
    xu[0] = d.alphas[bi] * EL(du, m, t, j);                           // isls.py:602-603
    _Pragma("unroll") for (int q = 1; q < C; q++) xu[q] = EL(Xu, m * C, t, j * C + q);
    _Pragma("unroll") for (int q = 0; q < C; q++) { z[q] = EL(Zm, m * C, t, j * C + q); lm[q] = EL(Lm, m * C, t, j * C + q); }
    _Pragma("unroll") for (int q = 0; q < C; q++) y[q] = (d.relax * xu[q] + (1.0 - d.relax) * z[q]) + lm[q];
    y[0] += un;                                                       // notebook cell 25
  }
  int its = 1;
  if (X.u_identity) {                                                 // isls_admm without project_u: z_u = y_u
    _Pragma("unroll") for (int q = 0; q < C; q++) zn[q] = y[q];
  } else {
    its = soc_project_set<CP, CC, CR>(S, S.b, y, zn, act, red);
  }
  double ps = 0.0, ds = 0.0;
  if (act) {
    zn[0] -= un;
    _Pragma("unroll") for (int q = 0; q < C; q++) {
      const double pr = xu[q] - zn[q], dz = zn[q] - z[q];
      lm[q] += pr;
      ps = fma(rho * pr, rho * pr, ps);
      ds = fma(rho * dz, rho * dz, ds);
      EL(Zm, m * C, t, j * C + q) = zn[q];
      EL(Lm, m * C, t, j * C + q) = lm[q];
    }
    EL(Xu, m * C, t, j * C) = xu[0];
    EL(rgu, m, t, j) = un + (zn[0] - lm[0]);                          // reg_abs of column 0 for the next k_ff
  }
  double prim = sqrt(block_sum(ps, red)), dual = sqrt(block_sum(ds, red));
  if (d.Zx) {
    // ---- state side (isls.py:631-638, 648-650): thread r = time step t, looping over the n rows (t, i) of
    // [d_x | Phi_x(:, :dim)];  z_x = project_x(alpha x_x + (1 - alpha) z_x + lambda_x, x_nom): column 0 shifted by
    // x_nom, the rows of every listed component projected by one project_set_convex call over that component's N rows
    // (its stop rule is the maximum over those rows), all other rows passed through; residuals weighted by Qr and
    // ADDED to the control side's
    constexpr int n = M::n;
    const bool actx = r < d.N;
    const int tx = actx ? r : 0;
    const double *xh = c.at(d.xh, d, n);
    double *Zx = c.at(d.Zx, d, n * C), *Lx = c.at(d.Lx, d, n * C), *Xx = c.at(d.Xx, d, n * C), *rgx = c.at(d.rgx, d, n);
    double psx = 0.0, dsx = 0.0;
    for (int ix = 0; ix < n; ix++) {
      int g = -1;
      for (int q = 0; q < X.ncomp; q++) if (X.comp[q] == ix) g = q;     // uniform over the CTA
      double xx[SOC_MAXC] = {}, zxv[SOC_MAXC] = {}, lxv[SOC_MAXC] = {}, yx[SOC_MAXC] = {}, znx[SOC_MAXC] = {};
      double xn0 = 0.0, rhx = 0.0;
      if (actx) {
        xn0 = EL(xh, n, tx, ix);
        rhx = d.rho_x[tx * n + ix];
        _Pragma("unroll") for (int q = 0; q < C; q++) {
          xx[q] = EL(Xx, n * C, tx, ix * C + q);
          zxv[q] = EL(Zx, n * C, tx, ix * C + q);
          lxv[q] = EL(Lx, n * C, tx, ix * C + q);
          yx[q] = (d.relax * xx[q] + (1.0 - d.relax) * zxv[q]) + lxv[q];
        }
        yx[0] += xn0;
      }
      if (g >= 0) {
        soc_project_set<CP, CC, CR>(S, X.b[g], yx, znx, actx, red);
      } else {
        _Pragma("unroll") for (int q = 0; q < C; q++) znx[q] = yx[q];
      }
      if (actx) {
        znx[0] -= xn0;
        _Pragma("unroll") for (int q = 0; q < C; q++) {
          const double pr = xx[q] - znx[q], dz = znx[q] - zxv[q];
          lxv[q] += pr;
          psx = fma(rhx * pr, rhx * pr, psx);
          dsx = fma(rhx * dz, rhx * dz, dsx);
          EL(Zx, n * C, tx, ix * C + q) = znx[q];
          EL(Lx, n * C, tx, ix * C + q) = lxv[q];
        }
        EL(rgx, n, tx, ix) = xn0 + (znx[0] - lxv[0]);                   // reg_abs of column 0 for the next k_ff
      }
    }
    prim = sqrt(block_sum(psx, red)) + prim;
    dual = sqrt(block_sum(dsx, red)) + dual;
  }
  if (r == 0) {
    if (d.out.inner_iters) d.out.inner_iters[((size_t)c.ob * d.max_outer + outer) * d.max_admm + inner] = its;
    d.cost_adm[c.b] = d.best_cost[c.b];                               // the search cost is the plain cost
    admm_finish<M>(d, c, outer, inner, bi, prim, dual);
  }
}

// d_u = x_u[:, 0] and Phi_u(:, :dim) = x_u[:, 1:] of the last ADMM iterate, natural layouts (isls.py:710-712)
template <class M>
__global__ void k_isls_out(Dev d, double *du_out, double *phi_out) {
  constexpr int m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (!c.valid) return;
  const int C = d.isls_C;
  const double *Xu = c.at(d.Xu, d, m * C);
  for (int t = 0; t < d.N; t++)
    for (int j = 0; j < m; j++) {
      du_out[((size_t)c.ob * d.N + t) * m + j] = EL(Xu, m * C, t, j * C);
      for (int q = 1; q < C; q++) phi_out[(((size_t)c.ob * d.N + t) * m + j) * (C - 1) + q - 1] = EL(Xu, m * C, t, j * C + q);
    }
}

// Obstacle-set projection of all rows of one problem + the rest of the ADMM update (admm.py:49-97): CTA = problem,
// thread = row t.  x = (x0 + rho sum_k (z_k - lambda_k)) / (1 + K rho); z_k = Pi_k(x + lambda_k); lambda_k += x - z_k;
// stop on the maximum over sets and rows of ||x - z_k||, rho ||z_k - z_k_prev|| (< threshold), or both maxima changing
// by < 1e-5 relative, or max_iter (projections.py:289-374 with As = I, bs = 0).
template <class M>
__global__ void k_obst_project(Dev d, int outer, int inner) {
  constexpr int n = M::n;
  __shared__ double red[32];
  const long long b = blockIdx.x;
  TileCtx<M> c(d, (int)(b / TILE), (int)(b % TILE));
  if (!c.valid || d.odone[c.b] || d.adone[c.b]) return;               // uniform over the CTA
  const int K = d.n_obst, N = d.N;
  const size_t arr = (size_t)d.T * N * n * TILE;
  const double *xw = c.at(d.obw, d, n), *pre = c.at(d.obw + arr, d, n);
  double *zx = c.at(d.zx, d, n), *lx = c.at(d.lx, d, n), *rgx = c.at(d.rgx, d, n);
  const double rho = d.obst_rho, inv = 1.0 / (1.0 + rho * K);         // inv(I + rho sum A_i'A_i), A_i = I
  double prx = 0.0, drx = 0.0;
  int its = 0;
  // rows are handled in slabs of blockDim.x; the stop rule couples all rows, so for N > blockDim.x every thread keeps
  // the state of its rows of every slab (N <= 1024 in all notebooks: one slab)
  const int t = threadIdx.x;
  const bool act = t < N;
  double x0[n], x[n], zk[ISLS_MAX_OBST][n], lk[ISLS_MAX_OBST][n];
#pragma unroll
  for (int i = 0; i < n; i++) {
    x0[i] = act ? EL(pre, n, t, i) : 0.0;
    x[i] = x0[i];
    for (int k = 0; k < K; k++) { zk[k][i] = x0[i]; lk[k][i] = 0.0; }  // z_i = A_i x0 + b_i, lambda_i = 0
  }
  double prim_ = 1e5, dual_ = 1e5;
  for (int j = 0; j < d.obst_max_iter; j++) {
    its = j + 1;
    double pmax = 0.0, dmax = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) {
      double r = 0.0;
      for (int k = 0; k < K; k++) r = r + (zk[k][i] - lk[k][i]);
      x[i] = inv * (x0[i] + rho * r);
    }
    for (int k = 0; k < K; k++) {
      double zn[n], y[2];
#pragma unroll
      for (int i = 0; i < n; i++) zn[i] = x[i] + lk[k][i];
      y[0] = zn[0]; y[1] = zn[1];
      obst_project_one(d, k, y);
      zn[0] = y[0]; zn[1] = y[1];
      double ps = 0.0, ds = 0.0;
#pragma unroll
      for (int i = 0; i < n; i++) {
        const double pr = x[i] - zn[i], du_ = rho * (zn[i] - zk[k][i]);
        ps += pr * pr;
        ds += du_ * du_;
        zk[k][i] = zn[i];
        lk[k][i] += pr;
      }
      pmax = fmax(pmax, sqrt(ps));
      dmax = fmax(dmax, sqrt(ds));
    }
    const double pprim = prim_, pdual = dual_;
    prim_ = block_max(act ? pmax : 0.0, red);
    dual_ = block_max(act ? dmax : 0.0, red);
    if (prim_ < d.obst_threshold && dual_ < d.obst_threshold) break;
    if (j < d.obst_max_iter - 1) {
      const double pc = fabs(pprim - prim_) / (pprim + 1e-30), dc = fabs(pdual - dual_) / (pdual + 1e-30);
      if (pc < 1e-5 && dc < 1e-5) break;
    }
  }
  if (act) {
#pragma unroll
    for (int i = 0; i < n; i++) {                                      // admm.py:49-59 with z = project_x(.)
      const double xv = EL(xw, n, t, i), zo = EL(zx, n, t, i);
      const double r = __dsub_rn(xv, x[i]), dz = __dsub_rn(x[i], zo);
      const double lv = __dadd_rn(EL(lx, n, t, i), r);
      prx = fma(r, r, prx);
      drx = fma(dz, dz, drx);
      EL(zx, n, t, i) = x[i];
      EL(lx, n, t, i) = lv;
      EL(rgx, n, t, i) = __dsub_rn(x[i], lv);
    }
  }
  const double ps = block_sum(prx, red), ds = block_sum(drx, red);
  if (threadIdx.x == 0) {
    if (d.out.inner_iters) d.out.inner_iters[((size_t)c.ob * d.max_outer + outer) * d.max_admm + inner] = its;
    const size_t S = (size_t)d.T * TILE;
    const double pru = d.cq[3 * S + c.b], dru = d.cq[4 * S + c.b];     // control-side sums parked by k_admm
    admm_finish<M>(d, c, outer, inner, d.best[c.b], sqrt(ps) + sqrt(pru), sqrt(ds) + sqrt(dru));
  }
}

// After ADMM (isls/isls.py:488-499): nominal <- last primal iterate, cost log, outer stop tests.
template <class M>
__global__ void k_outer_end(Dev d, int outer) {
  constexpr int n = M::n, m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b]) return;
  // nominal <- last primal iterate (isls.py:488): re-roll u^ + alpha* du in place (same arithmetic as k_admm, whose
  // cost of this iterate is cost_adm) instead of storing every ADMM iterate's trajectory
  double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *du = c.at(d.du, d, m);
  {
    const double al = d.alphas[d.best[c.b]];
    double x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, 0, i);
    constexpr int UB = 8;                          // steps of loads in flight before the first store
    for (int t0 = 0; t0 < d.N; t0 += UB) {
      // branch-free load phase (t clamped, raw operands): the guarded form with the FMA inside compiled to one
      // load-then-use block per step, i.e. one memory round trip per step (see the line-search epilogue)
      double dv[UB][m], hv[UB][m];
#pragma unroll
      for (int q = 0; q < UB; q++) {
        const int t = min(t0 + q, d.N - 1);
#pragma unroll
        for (int j = 0; j < m; j++) { dv[q][j] = EL(du, m, t, j); hv[q][j] = EL(uh, m, t, j); }
      }
#pragma unroll
      for (int q = 0; q < UB; q++) {
        const int t = t0 + q;
        if (t < d.N) {
#pragma unroll
          for (int j = 0; j < m; j++) { u[j] = fma(al, dv[q][j], hv[q][j]); EL(uh, m, t, j) = u[j]; }
#pragma unroll
          for (int i = 0; i < n; i++) EL(xh, n, t, i) = x[i];
          M::step(x, u, xn, d.dt);
#pragma unroll
          for (int i = 0; i < n; i++) x[i] = xn[i];
        }
      }
    }
  }
  const double cost = d.cost_adm[c.b], prev = d.prev_cost[c.b];
  d.cost[c.b] = cost;
  const int nl = d.nlog[c.b];
  d.nlog[c.b] = nl + 1;
  d.oit[c.b] = outer + 1;
  double *cl = d.out.cost_log + c.ob * (d.max_outer + 1);
  if (c.valid) cl[nl] = cost;
  if (d.fixed_budget) return;
  int st = 0;
  if (fabs(cost - prev) < d.outer_tol) st = ISLS_ST_CONVERGED_COST;              // isls.py:493
  else {
    // |mean(cost_log[-4:]) - mean(cost_log[-8:-4])| < 1e-3 with python slice semantics (isls.py:497)
    const int len = nl + 1;
    const int a0 = max(0, len - 4), p0 = max(0, len - 8), p1 = max(0, len - 4);
    if (p1 > p0) {
      double s1 = 0.0, s2 = 0.0;
      for (int i = a0; i < len; i++) s1 += (i == nl) ? cost : cl[i];
      for (int i = p0; i < p1; i++) s2 += cl[i];
      if (fabs(s1 / (len - a0) - s2 / (p1 - p0)) < d.osc_tol) st = ISLS_ST_OSCILLATING;
    }
  }
  if (st) {
    d.status[c.b] |= st;
    d.odone[c.b] = 1;
    if (d.orig && c.valid) retire<M>(d, c, true);      // compacting mode: results leave the workspace now
  }
}

// Write one problem's results to the natural (reference) layouts.
template <class M>
__device__ __forceinline__ void retire(const Dev &d, const TileCtx<M> &c, bool finished) {
  constexpr int n = M::n, m = M::m;
  const isls_solve_out &o = d.out;
  auto unpack = [&](const double *src, double *dst, int dim) {
    if (!dst) return;
    const double *s = c.at(const_cast<double *>(src), d, dim);
    double *q = dst + (size_t)c.ob * d.N * dim;
    const int tot = d.N * dim;                    // natural layout [t][i] == tile layout row index t*dim + i
    int r = 0;
    for (; r + 8 <= tot; r += 8) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = s[(size_t)(r + k) * TILE];
#pragma unroll
      for (int k = 0; k < 8; k++) q[r + k] = v[k];
    }
    for (; r < tot; r++) q[r] = s[(size_t)r * TILE];
  };
  unpack(d.xh, o.x, n);
  unpack(d.uh, o.u, m);
  unpack(d.zx, o.z_x, n);
  unpack(d.zu, o.z_u, m);
  unpack(d.lx, o.lam_x, n);
  unpack(d.lu, o.lam_u, m);
  unpack(d.Kg, o.K, m * n);
  unpack(d.kk, o.k, m);
  if (o.Qux || o.Quu || o.Quu_inv) {            // Riccati logs of the last K-pass (sls.py:159-162); row N-1 is zero
    constexpr int nt = NTRI(M::m);
    const double *qx = c.at(d.Qux, d, m * n), *qu = c.at(d.Quu, d, nt), *qi = c.at(d.Qui, d, nt);
    for (int t = 0; t < d.N; t++) {
      const bool last = t == d.N - 1;
      if (o.Qux)
        for (int q = 0; q < m * n; q++) o.Qux[((size_t)c.ob * d.N + t) * m * n + q] = last ? 0.0 : EL(qx, m * n, t, q);
      for (int a = 0; a < m; a++)
        for (int b2 = 0; b2 < m; b2++) {
          const int q = a >= b2 ? tri(a, b2) : tri(b2, a);
          if (o.Quu) o.Quu[(((size_t)c.ob * d.N + t) * m + a) * m + b2] = last ? 0.0 : EL(qu, nt, t, q);
          if (o.Quu_inv) o.Quu_inv[(((size_t)c.ob * d.N + t) * m + a) * m + b2] = last ? 0.0 : EL(qi, nt, t, q);
        }
    }
  }
  int st = d.status[c.b];
  if (!finished) st |= ISLS_ST_MAX_ITER;
  if (o.cost) o.cost[c.ob] = d.cost[c.b];
  if (o.status) o.status[c.ob] = st;
  if (o.n_log) o.n_log[c.ob] = d.nlog[c.b];
  if (o.outer_iters) o.outer_iters[c.ob] = d.oit[c.b];
}

// Unpack results to the natural (reference) layouts.
template <class M>
__global__ void k_finalize(Dev d) {
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (!c.valid) return;
  retire<M>(d, c, d.odone[c.b] != 0);
}

// compaction scan (model-independent kernel, defined in isls_b200.cu)
void isls_launch_compact_scan(int nslots, const int *odone, const int *orig, int *newpos, int *nact, cudaStream_t s);

template <class M>
__global__ void k_compact_move(Dev a, Dev b) {           // a: current buffers, b: alternate buffers
  constexpr int n = M::n, m = M::m;
  const int tile = blockIdx.x * blockDim.y + threadIdx.y;
  if (tile >= a.T) return;
  const int lane = threadIdx.x;
  const long long s = (long long)tile * TILE + lane;
  const int np = a.newpos[s];
  const int na = a.nact[0];
  // slots at or beyond the active count become empty in the new buffers
  if (s >= na) { b.odone[s] = 1; b.orig[s] = -1; b.status[s] = 0; }
  if (np < 0) return;
  const int nt_ = np / TILE, nl_ = np % TILE;
  auto mv = [&](const double *src, double *dst, int rows, int dim) {
    const double *p = src + (size_t)tile * rows * dim * TILE + lane;
    double *q = dst + (size_t)nt_ * rows * dim * TILE + nl_;
    const int tot = rows * dim;
    int r = 0;
    for (; r + 8 <= tot; r += 8) {               // 8 loads in flight before the stores
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = p[(size_t)(r + k) * TILE];
#pragma unroll
      for (int k = 0; k < 8; k++) q[(size_t)(r + k) * TILE] = v[k];
    }
    for (; r < tot; r++) q[(size_t)r * TILE] = p[(size_t)r * TILE];
  };
  mv(a.xh, b.xh, a.N, n);
  mv(a.uh, b.uh, a.N, m);
  mv(a.zx, b.zx, a.N, n);
  mv(a.zu, b.zu, a.N, m);
  mv(a.zs, b.zs, a.n_via, n);
  b.cost[np] = a.cost[s];
  b.nlog[np] = a.nlog[s];
  b.status[np] = a.status[s];
  b.oit[np] = a.oit[s];
  b.orig[np] = a.orig[s];
  b.odone[np] = 0;
  b.adone[np] = 0;
}

// ------------------------------------------------------------------------------------------- plain iLQR (DP) kernels
// Full backward pass of iSLS.backward_pass_DP (isls/isls.py:229-308, `Cts is None` branch): K_t and k_t with
// cx = 2Q(x^-z), cu = 2R u^; K[N-1] = k[N-1] = 0.
template <class M>
__global__ void k_backward_full(Dev d) {
  constexpr int n = M::n, m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b]) return;
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  double *Kg = c.at(d.Kg, d, m * n), *kk = c.at(d.kk, d, m);
  double A[n][n], Bm[n][m], V[n][n], v[n];
  init_AB<M>(A, Bm);
  {
    const int t = d.N - 1;
    double xl[n], hl[n];
#pragma unroll
    for (int i = 0; i < n; i++) xl[i] = EL(xh, n, t, i);
    state_grad_hess<M>(d, zs, t, xl, v, hl);                                                       // isls.py:252 / 258
#pragma unroll
    for (int i = 0; i < n; i++) {
#pragma unroll
      for (int j = 0; j < n; j++) V[i][j] = (i == j) ? hl[i] : 0.0;                                // isls.py:251 / 257
    }
#pragma unroll
    for (int q = 0; q < m * n; q++) EL(Kg, m * n, t, q) = 0.0;
#pragma unroll
    for (int j = 0; j < m; j++) EL(kk, m, t, j) = 0.0;
  }
  bool ok = true;
  for (int t = d.N - 2; t >= 0; t--) {
    double x[n], u[m], J[M::NJA], dxx[n], duu[m], cx[n], cu[m], K[m][n], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, t, i);
    state_grad_hess<M>(d, zs, t, x, cx, dxx);
#pragma unroll
    for (int j = 0; j < m; j++) {
      u[j] = EL(uh, m, t, j);
      duu[j] = 2.0 * (d.u_std * d.Rw[j]);
      cu[j] = 2.0 * (d.u_std * d.Rw[j]) * u[j];
    }
    M::jac(x, u, J, d.dt);
    M::expand(J, A, Bm, d.dt);
    // the feed-forward step needs the pre-update V only through K,Qux,Quu of this step and the old v
    double Vn[n][n];
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
      for (int j = 0; j < n; j++) Vn[i][j] = V[i][j];
    ok &= riccati_step<M>(A, Bm, dxx, duu, Vn, K, Qux, Quu, Qui);
    ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
      for (int j = 0; j < n; j++) V[i][j] = Vn[i][j];
#pragma unroll
    for (int a = 0; a < m; a++) {
#pragma unroll
      for (int j = 0; j < n; j++) EL(Kg, m * n, t, a * n + j) = K[a][j];
      EL(kk, m, t, a) = kt[a];
    }
  }
  if (!ok) d.status[c.b] |= ISLS_ST_NON_PD;
}

// Closed-loop line search (isls/isls.py:310-334, 357-363): u = K(x - x^) + alpha k + u^, NaN cost -> 1e5, argmin.
template <class M, int CPT, int MAXW, int MINB>
__global__ void __launch_bounds__(TILE * MAXW, MINB) k_linesearch_closed(Dev d) {
  constexpr int n = M::n, m = M::m;
  __shared__ double sc[MAX_L][TILE];
  const int tile = d.tile0 + blockIdx.x;
  TileCtx<M> c(d, tile, threadIdx.x);
  const int w = threadIdx.y;
  const bool skip = d.odone[c.b];
  if (__syncthreads_and(skip)) return;
  if (!skip) {
    const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m), *kk = c.at(d.kk, d, m);
    const double *Kg = c.at(d.Kg, d, m * n);
    const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
    double al[CPT], x[CPT][n], cs[CPT], cc[CPT];
#pragma unroll
    for (int q = 0; q < CPT; q++) {
      const int l = w * CPT + q;
      al[q] = l < d.L ? d.alphas[l] : 0.0;
      cs[q] = cc[q] = 0.0;
#pragma unroll
      for (int i = 0; i < n; i++) x[q][i] = EL(xh, n, 0, i);
    }
    for (int t = 0; t < d.N; t++) {
      double xn0[n], un[m], kt[m], K[m][n], zv[n], qd[n];
      const bool qz = d.qnz[t];
#pragma unroll
      for (int i = 0; i < n; i++) xn0[i] = EL(xh, n, t, i);
#pragma unroll
      for (int j = 0; j < m; j++) {
        un[j] = EL(uh, m, t, j);
        kt[j] = EL(kk, m, t, j);
#pragma unroll
        for (int i = 0; i < n; i++) K[j][i] = EL(Kg, m * n, t, j * n + i);
      }
      if (qz) {
        const int s = d.seq[t];
#pragma unroll
        for (int i = 0; i < n; i++) { zv[i] = EL(zs, n, s, i); qd[i] = d.qd[t * n + i]; }
      }
#pragma unroll
      for (int q = 0; q < CPT; q++) {
        double u[m], xn[n];
#pragma unroll
        for (int j = 0; j < m; j++) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < n; i++) acc = fma(K[j][i], x[q][i] - xn0[i], acc);
          u[j] = (acc + al[q] * kt[j]) + un[j];                       // isls.py:329
          cc[q] += (d.Rw[j] * u[j]) * u[j];
        }
        if (qz) {
          if (d.cost_kind == ISLS_COST_QUADRATIC) {
#pragma unroll
            for (int i = 0; i < n; i++) { const double e = x[q][i] - zv[i]; cs[q] += (e * e) * qd[i]; }
          } else cs[q] += state_cost<M>(d, zs, t, x[q]);
        }
        M::step(x[q], u, xn, d.dt);
#pragma unroll
        for (int i = 0; i < n; i++) x[q][i] = xn[i];
      }
    }
#pragma unroll
    for (int q = 0; q < CPT; q++) {
      const int l = w * CPT + q;
      if (l < d.L) {
        double tot = cs[q] + d.u_std * cc[q];
        if (tot != tot) {
          // isls.py:362 (benign race: same bit).  The Tutorial's cost closure maps NaN to 1e6 itself (cell 14), so
          // with that cost the solver never sees a NaN and no flag is raised.
          if (d.cost_kind == ISLS_COST_QUADRATIC) { tot = 1e5; d.status[c.b] |= ISLS_ST_NAN_COST; }
          else tot = 1e6;
        }
        sc[l][c.lane] = tot;
      }
    }
  }
  __syncthreads();
  if (w == 0 && !skip) {
    bool has_nan;
    const int idx = argmin_np(&sc[0][c.lane], d.L, TILE, &has_nan);
    d.best[c.b] = idx;
    d.best_cost[c.b] = sc[idx][c.lane];
  }
}

// Accept test + nominal update + stop rules of plain iLQR (isls/isls.py:364-370, 125-131).
template <class M>
__global__ void k_accept_closed(Dev d, int it) {
  constexpr int n = M::n, m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (d.odone[c.b]) return;
  double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m);
  const double *kk = c.at(d.kk, d, m), *Kg = c.at(d.Kg, d, m * n);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  const int bi = d.best[c.b];
  const bool nonpd = d.status[c.b] & ISLS_ST_NON_PD;
  const bool ok = (d.best_cost[c.b] - d.cost[c.b] < 0.0) && !nonpd;                 // isls.py:365-367
  d.oit[c.b] = it + 1;
  if (c.valid && d.out.alpha_idx) d.out.alpha_idx[(size_t)c.ob * d.max_outer + it] = ok ? bi : -1;
  int nl = d.nlog[c.b];
  double *cl = d.out.cost_log + c.ob * (d.max_outer + 1);
  double newc = d.cost[c.b];
  if (ok) {
    const double al = d.alphas[bi];
    double x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = EL(xh, n, 0, i);
    double cs = 0.0, cc = 0.0;
    for (int t = 0; t < d.N; t++) {
#pragma unroll
      for (int j = 0; j < m; j++) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) acc = fma(EL(Kg, m * n, t, j * n + i), x[i] - EL(xh, n, t, i), acc);
        u[j] = (acc + al * EL(kk, m, t, j)) + EL(uh, m, t, j);
        cc += (d.Rw[j] * u[j]) * u[j];
      }
#pragma unroll
      for (int j = 0; j < m; j++) EL(uh, m, t, j) = u[j];
#pragma unroll
      for (int i = 0; i < n; i++) EL(xh, n, t, i) = x[i];
      cs += state_cost<M>(d, zs, t, x);
      M::step(x, u, xn, d.dt);
#pragma unroll
      for (int i = 0; i < n; i++) x[i] = xn[i];
    }
    newc = cs + d.u_std * cc;
    d.cost[c.b] = newc;
    if (c.valid) cl[nl] = newc;
    nl += 1;
    d.nlog[c.b] = nl;
  }
  if (d.fixed_budget) return;
  int st = 0;
  if (nl >= 2) {
    const double last = ok ? newc : cl[nl - 1];
    if (fabs(last - cl[nl - 2]) < d.tol) st = ISLS_ST_CONVERGED_COST;                // isls.py:125
  }
  if (!st && !ok) st = ISLS_ST_LINESEARCH_FAIL;                                      // isls.py:128
  if (st) {
    d.status[c.b] |= st;
    d.odone[c.b] = 1;
  }
}

// -------------------------------------------------------------------------------------------- stage-level kernels
template <class M>
__global__ void k_pack_stage(Dev d, const double *x_nom, const double *u_nom, const double *du_in,
                             const double *zs_in, const double *regx, const double *regu) {
  constexpr int n = M::n, m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m), *du = c.at(d.du, d, m);
  double *rgx = c.at(d.rgx, d, n), *rgu = c.at(d.rgu, d, m);
  double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  for (int k = 0; k < d.n_via; k++)
    for (int i = 0; i < n; i++) EL(zs, n, k, i) = zs_in[(c.ob * d.n_via + k) * n + i];
  for (int t = 0; t < d.N; t++) {
    for (int i = 0; i < n; i++) {
      EL(xh, n, t, i) = x_nom[(c.ob * d.N + t) * n + i];
      if (d.proj_x) EL(rgx, n, t, i) = regx[(c.ob * d.N + t) * n + i];
    }
    for (int j = 0; j < m; j++) {
      EL(uh, m, t, j) = u_nom[(c.ob * d.N + t) * m + j];
      EL(du, m, t, j) = du_in[(c.ob * d.N + t) * m + j];
      if (d.proj_u) EL(rgu, m, t, j) = regu[(c.ob * d.N + t) * m + j];
    }
  }
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
  for (int t = 0; t < d.N; t++)
    for (int j = 0; j < m; j++) {
      const double u = EL(uh, m, t, j), dv = EL(du, m, t, j);
      r0 = fma(d.Rw[j] * u, u, r0);
      r1 = fma(d.Rw[j] * u, dv, r1);
      r2 = fma(d.Rw[j] * dv, dv, r2);
      if (d.proj_u) {
        const double rho = d.rho_u[t * m + j], e = u - EL(rgu, m, t, j);
        c0 = fma(rho * e, e, c0);
        c1 = fma(2.0 * rho * e, dv, c1);
        c2 = fma(rho * dv, dv, c2);
      }
    }
  const size_t S = (size_t)d.T * TILE;
  r0 *= d.u_std;
  r1 *= 2.0 * d.u_std;
  r2 *= d.u_std;
  d.cq[c.b] = c0 + r0;
  d.cq[S + c.b] = c1 + r1;
  d.cq[2 * S + c.b] = c2 + r2;
  d.cq[3 * S + c.b] = r0;
  d.cq[4 * S + c.b] = r1;
  d.cq[5 * S + c.b] = r2;
  d.odone[c.b] = 0;
  d.adone[c.b] = 0;
  d.status[c.b] = 0;
}

template <class M>
__global__ void k_unpack_stage(Dev d, double *costs, int *best, double *x_best, double *u_best) {
  constexpr int n = M::n, m = M::m;
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  if (!c.valid) return;
  // re-roll the winner (rollout_batch, isls/isls.py:135-154)
  const double *xh = c.at(d.xh, d, n), *uh = c.at(d.uh, d, m), *du = c.at(d.du, d, m);
  const int bi = d.best[c.b];
  const double al = d.alphas[bi];
  best[c.b] = bi;
  const double *lsc = d.lsc + (size_t)tile * d.L * TILE + c.lane;
  for (int l = 0; l < d.L; l++) costs[c.b * d.L + l] = lsc[(size_t)l * TILE];
  double x[n], u[m], xn[n];
  for (int i = 0; i < n; i++) x[i] = EL(xh, n, 0, i);
  for (int t = 0; t < d.N; t++) {
    for (int j = 0; j < m; j++) {
      u[j] = EL(uh, m, t, j) + al * EL(du, m, t, j);
      u_best[(c.b * d.N + t) * m + j] = u[j];
    }
    for (int i = 0; i < n; i++) x_best[(c.b * d.N + t) * n + i] = x[i];
    M::step(x, u, xn, d.dt);
    for (int i = 0; i < n; i++) x[i] = xn[i];
  }
}

// State projection of the spherical-obstacle notebook on the LQT path (one problem per thread, rows swept sequentially;
// Double integrator/LQR and SLS with spherical obstacle avoidance.ipynb cell 12): positions through project_set_convex
// (As = I_2, bs = 0; projections.py:289-374), then project_set_convex_dykstra (projections.py:465-505); the other state
// components pass through.  obw arrays [T][N][n][32]: 0 pre-projection point (in), 1 result (out), 2.. scratch
// (z_k, lambda_k of the consensus ADMM in components 0-1 / 2-3 of array 2 + k; Dykstra increments in array 2 + K + k).
template <class M>
__device__ __forceinline__ void lqt_obst_project(const Dev &d, const TileCtx<M> &c, int *its1, int *its2) {
  constexpr int n = M::n;
  const int K = d.n_obst, N = d.N;
  const size_t arr = (size_t)d.T * N * n * TILE;
  const double *pre = c.at(d.obw, d, n);
  double *res = c.at(d.obw + arr, d, n);
  const double rho = d.obst_rho, inv = 1.0 / (1.0 + rho * K);
  double prim_ = 1e5, dual_ = 1e5;
  int it = 0;
  for (int j = 0; j < d.obst_max_iter; j++) {
    it = j + 1;
    double pmax = 0.0, dmax = 0.0;
    for (int t = 0; t < N; t++) {
      double x0[2], x[2], zk[ISLS_MAX_OBST][2], lk[ISLS_MAX_OBST][2];
      x0[0] = EL(pre, n, t, 0); x0[1] = EL(pre, n, t, 1);
      for (int k = 0; k < K; k++) {
        const double *sk = c.at(d.obw + (2 + k) * arr, d, n);
        for (int i = 0; i < 2; i++) { zk[k][i] = j == 0 ? x0[i] : EL(sk, n, t, i); lk[k][i] = j == 0 ? 0.0 : EL(sk, n, t, 2 + i); }
      }
      for (int i = 0; i < 2; i++) {
        double r = 0.0;
        for (int k = 0; k < K; k++) r = r + (zk[k][i] - lk[k][i]);
        x[i] = inv * (x0[i] + rho * r);
        EL(res, n, t, i) = x[i];
      }
      for (int k = 0; k < K; k++) {
        double *sk = c.at(d.obw + (2 + k) * arr, d, n);
        double y[2] = {x[0] + lk[k][0], x[1] + lk[k][1]};
        obst_project_one(d, k, y);
        double ps = 0.0, ds = 0.0;
        for (int i = 0; i < 2; i++) {
          const double pr = x[i] - y[i], du_ = rho * (y[i] - zk[k][i]);
          ps += pr * pr;
          ds += du_ * du_;
          EL(sk, n, t, i) = y[i];
          EL(sk, n, t, 2 + i) = lk[k][i] + pr;
        }
        pmax = fmax(pmax, sqrt(ps));
        dmax = fmax(dmax, sqrt(ds));
      }
    }
    const double pprim = prim_, pdual = dual_;
    prim_ = pmax;
    dual_ = dmax;
    if (prim_ < d.obst_threshold && dual_ < d.obst_threshold) break;
    if (j < d.obst_max_iter - 1) {
      const double pc = fabs(pprim - prim_) / (pprim + 1e-30), dc = fabs(pdual - dual_) / (pdual + 1e-30);
      if (pc < 1e-5 && dc < 1e-5) break;
    }
  }
  *its1 = it;
  // ---- Dykstra: u = x; z_i = 0; sweep the sets while any row's summed squared increment change is >= tol
  int kd = 0;
  if (d.obst_dyk_max_iter > 0) {
    bool any = true;
    while (kd <= d.obst_dyk_max_iter && any) {
      any = false;
      for (int t = 0; t < N; t++) {
        double u[2] = {EL(res, n, t, 0), EL(res, n, t, 1)};
        double cI = 0.0;
        for (int k = 0; k < K; k++) {
          double *zd = c.at(d.obw + (2 + K + k) * arr, d, n);
          const double pz0 = kd == 0 ? 0.0 : EL(zd, n, t, 0), pz1 = kd == 0 ? 0.0 : EL(zd, n, t, 1);
          const double pu0 = u[0], pu1 = u[1];
          double y[2] = {pu0 - pz0, pu1 - pz1};
          obst_project_one(d, k, y);
          u[0] = y[0]; u[1] = y[1];
          const double nz0 = u[0] - (pu0 - pz0), nz1 = u[1] - (pu1 - pz1);
          EL(zd, n, t, 0) = nz0; EL(zd, n, t, 1) = nz1;
          const double e0 = pz0 - nz0, e1 = pz1 - nz1, nr = sqrt(e0 * e0 + e1 * e1);
          cI += nr * nr;                                        // np.linalg.norm(.)**2
        }
        EL(res, n, t, 0) = u[0]; EL(res, n, t, 1) = u[1];
        any |= cI >= d.obst_dyk_tol;
      }
      kd++;
    }
  }
  *its2 = kd;
  for (int t = 0; t < N; t++)
#pragma unroll
    for (int i = 2; i < n; i++) EL(res, n, t, i) = EL(pre, n, t, i);
}

// ------------------------------------------------------------------------------------------- LQT-ADMM (DP) kernel
// SLS.ADMM_LQT_DP (isls/sls.py:298-317): the gains K, Quu, Quu^-1, Qux depend only on (A, B, Q, rho), so one
// thread block computes them once (k_kpass on a single tile with a zero nominal works for linear models because
// the Jacobian does not depend on the trajectory); every problem then iterates ff-pass (sls.py:168-202) +
// closed-loop rollout from its x0 (sls_base.py:76-89) + projection / dual update (admm.py) inside ONE kernel.
// Everything that is the same for all problems and all iterations - the shared gains K, Qux, Quu, Quu^-1 and the plan
// constants (Q, rho, bounds, seq) - is staged ONCE into shared memory when it fits (SM = true: 18 KB at N = 50, n = 4,
// m = 2), and each thread's own operands (z, lambda, k) are fetched one step ahead into registers: a problem's ADMM
// loop is a chain of max_iter x 2N dependent steps run by a lone warp, and with plain loads every step paid an L2 round
// trip (C1: 90 us per iteration; profiles/r1_c3_small_batch_kernels.md).
template <class M, bool SM>
__global__ void k_lqt_admm(Dev d_in, const double *x0_in) {
  constexpr int n = M::n, m = M::m;
  constexpr int nt = NTRI(M::m);
  extern __shared__ double smem_lqt[];
  Dev d = d_in;
  const int N_ = d.N;
  const double *sK = nullptr, *sQx = nullptr, *sQu = nullptr, *sQi = nullptr;
  if (SM) {
    double *w = smem_lqt;
    const int nthr = blockDim.x * blockDim.y, tid0 = threadIdx.y * blockDim.x + threadIdx.x;
    auto stage_gain = [&](const double *g, int dim) -> const double * {     // tile 0 / lane 0 of a tile-blocked array
      double *dst = w;
      for (int q = tid0; q < N_ * dim; q += nthr) dst[q] = g[(size_t)q * TILE];
      w += N_ * dim;
      return dst;
    };
    auto stage_plan = [&](const double *g, int dim) -> const double * {
      double *dst = w;
      for (int q = tid0; q < N_ * dim; q += nthr) dst[q] = g ? g[q] : 0.0;
      w += N_ * dim;
      return g ? dst : nullptr;
    };
    sK = stage_gain(d.Kg, m * n);
    sQx = stage_gain(d.Qux, m * n);
    sQu = stage_gain(d.Quu, nt);
    sQi = stage_gain(d.Qui, nt);
    d.qd = stage_plan(d_in.qd, n);
    d.rho_x = stage_plan(d_in.rho_x, n);
    d.lo_x = stage_plan(d_in.lo_x, n);
    d.hi_x = stage_plan(d_in.hi_x, n);
    d.rho_u = stage_plan(d_in.rho_u, m);
    d.lo_u = stage_plan(d_in.lo_u, m);
    d.hi_u = stage_plan(d_in.hi_u, m);
    int *wi = reinterpret_cast<int *>(w);
    for (int q = tid0; q < N_; q += nthr) { wi[q] = d_in.seq[q]; wi[N_ + q] = d_in.qnz[q]; }
    d.seq = wi;
    d.qnz = wi + N_;
    __syncthreads();
  }
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  // gains are shared: tile 0 / lane 0 of the gain arrays
  const double *Kg = d.Kg, *Qx = d.Qux, *Qu = d.Quu, *Qi = d.Qui;
  auto gK = [&](int t, int q) -> double { return SM ? sK[t * (m * n) + q] : EL(Kg, m * n, t, q); };
  auto gQx = [&](int t, int q) -> double { return SM ? sQx[t * (m * n) + q] : EL(Qx, m * n, t, q); };
  auto gQu = [&](int t, int q) -> double { return SM ? sQu[t * nt + q] : EL(Qu, nt, t, q); };
  auto gQi = [&](int t, int q) -> double { return SM ? sQi[t * nt + q] : EL(Qi, nt, t, q); };
  double *xa = c.at(d.xh, d, n), *ua = c.at(d.uh, d, m), *kk = c.at(d.kk, d, m);   // primal iterate -> result
  double *zx = c.at(d.zx, d, n), *lx = c.at(d.lx, d, n);
  double *zu = c.at(d.zu, d, m), *lu = c.at(d.lu, d, m);
  const double *zs = d.zs + (size_t)tile * d.n_via * n * TILE + c.lane;
  double A[n][n], Bm[n][m], J[M::NJA];
  init_AB<M>(A, Bm);
  M::expand(J, A, Bm, d.dt);
  double x0[n];
#pragma unroll
  for (int i = 0; i < n; i++) x0[i] = x0_in[c.ob * n + i];
  double prim = 1e6, dual = 1e6;
  int ex = 0, it = 0;
  int8_t *mkx = (d.out.mask_x && c.valid) ? d.out.mask_x + c.ob * d.N * n : nullptr;
  int8_t *mku = (d.out.mask_u && c.valid) ? d.out.mask_u + c.ob * d.N * m : nullptr;
  double cost = 0.0;
  for (it = 0; it < d.max_admm && !ex; it++) {
    // ---- ff-pass with cx = -2Q z_via - 2Qr reg_x, cu = -2Rr reg_u   (sls.py:187-193, absolute coordinates)
    double v[n];
    {
      const int t = d.N - 1, s = d.seq[t];
#pragma unroll
      for (int i = 0; i < n; i++) {
        double g = d.qnz[t] ? -2.0 * d.qd[t * n + i] * EL(zs, n, s, i) : 0.0;
        if (d.proj_x) g += -2.0 * d.rho_x[t * n + i] * (EL(zx, n, t, i) - EL(lx, n, t, i));
        v[i] = g;
      }
#pragma unroll
      for (int j = 0; j < m; j++) {
        // DP form: k[N-1] = 0 (sls.py:113-114); batch form (ADMM_LQT_Batch, sls.py:283-286): the last control is
        // solved for, u_{N-1} = (R + Rr)^-1 Rr reg_u
        const double cuL = d.proj_u ? -2.0 * d.rho_u[t * m + j] * (EL(zu, m, t, j) - EL(lu, m, t, j)) : 0.0;
        EL(kk, m, t, j) = d.last_stage_dp ? 0.0 : -cuL / (2.0 * (d.u_std * d.Rw[j] + d.rho_u[t * m + j]));
      }
    }
    // this thread's operands of the next step to process, fetched one step ahead
    double pzx[n], plx[n], pzu[m], plu[m], pkk[m];
    auto fetch_own = [&](int t, bool with_k) {
      if (d.proj_x) {
#pragma unroll
        for (int i = 0; i < n; i++) { pzx[i] = EL(zx, n, t, i); plx[i] = EL(lx, n, t, i); }
      }
#pragma unroll
      for (int j = 0; j < m; j++) {
        if (d.proj_u) { pzu[j] = EL(zu, m, t, j); plu[j] = EL(lu, m, t, j); }
        if (with_k) pkk[j] = EL(kk, m, t, j);
      }
    };
    if (d.N >= 2) fetch_own(d.N - 2, false);
    for (int t = d.N - 2; t >= 0; t--) {
      double cx[n], cu[m], Qux[m][n], Quu[m][m], Qui[m][m], kt[m];
      double czx[n], clx[n], czu[m], clu[m];
#pragma unroll
      for (int i = 0; i < n; i++) { czx[i] = pzx[i]; clx[i] = plx[i]; }
#pragma unroll
      for (int j = 0; j < m; j++) { czu[j] = pzu[j]; clu[j] = plu[j]; }
      if (t > 0) fetch_own(t - 1, false);
      const int s = d.seq[t];
#pragma unroll
      for (int i = 0; i < n; i++) {
        double g = d.qnz[t] ? -2.0 * d.qd[t * n + i] * EL(zs, n, s, i) : 0.0;
        if (d.proj_x) g += -2.0 * d.rho_x[t * n + i] * (czx[i] - clx[i]);
        cx[i] = g;
      }
#pragma unroll
      for (int j = 0; j < m; j++)
        cu[j] = d.proj_u ? -2.0 * d.rho_u[t * m + j] * (czu[j] - clu[j]) : 0.0;
#pragma unroll
      for (int a = 0; a < m; a++) {
#pragma unroll
        for (int j = 0; j < n; j++) Qux[a][j] = gQx(t, a * n + j);
#pragma unroll
        for (int b2 = 0; b2 <= a; b2++) {
          Qui[a][b2] = gQi(t, tri(a, b2)); Qui[b2][a] = Qui[a][b2];
          Quu[a][b2] = gQu(t, tri(a, b2)); Quu[b2][a] = Quu[a][b2];
        }
      }
      ff_step<M>(A, Bm, cx, cu, Qux, Quu, Qui, v, kt);
#pragma unroll
      for (int j = 0; j < m; j++) EL(kk, m, t, j) = kt[j];
    }
    // ---- closed-loop rollout u = K x + k (sls_base.py:76-89) fused with the ADMM update (admm.py:43-69)
    double x[n], u[m], xn[n];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = x0[i];
    double cs = 0.0, cc = 0.0, prx = 0.0, pru = 0.0, drx = 0.0, dru = 0.0;
    fetch_own(0, true);
    for (int t = 0; t < d.N; t++) {
      // the step's own operands were fetched during the previous step; the next step's go out before this step's stores
      double zuv[m], luv[m], kv[m], zxv[n], lxv[n], Kt[m][n];
#pragma unroll
      for (int j = 0; j < m; j++) {
#pragma unroll
        for (int i = 0; i < n; i++) Kt[j][i] = gK(t, j * n + i);
        kv[j] = pkk[j];
        zuv[j] = pzu[j];
        luv[j] = plu[j];
      }
#pragma unroll
      for (int i = 0; i < n; i++) { zxv[i] = pzx[i]; lxv[i] = plx[i]; }
      if (t + 1 < d.N) fetch_own(t + 1, true);
      if (d.n_obst > 0) {          // park the pre-projection point; the rows are projected together after the rollout
        double *pre = c.at(d.obw, d, n);
#pragma unroll
        for (int i = 0; i < n; i++)
          EL(pre, n, t, i) = __dadd_rn(__dadd_rn(__dmul_rn(d.relax, x[i]), __dmul_rn(__dsub_rn(1.0, d.relax), zxv[i])), lxv[i]);
      }
#pragma unroll
      for (int j = 0; j < m; j++) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) acc = fma(Kt[j][i], x[i], acc);
        u[j] = acc + kv[j];
        EL(ua, m, t, j) = u[j];
        cc += (d.Rw[j] * u[j]) * u[j];
        if (d.proj_u) {
          int mk;
          admm_elem(u[j], d.relax, d.lo_u[t * m + j], d.hi_u[t * m + j], zuv[j], luv[j], pru, dru, mk);
          EL(zu, m, t, j) = zuv[j];
          EL(lu, m, t, j) = luv[j];
          if (mku) mku[t * m + j] = (int8_t)mk;
        }
      }
#pragma unroll
      for (int i = 0; i < n; i++) {
        EL(xa, n, t, i) = x[i];
        if (d.proj_x && d.n_obst == 0) {
          int mk;
          admm_elem(x[i], d.relax, d.lo_x[t * n + i], d.hi_x[t * n + i], zxv[i], lxv[i], prx, drx, mk);
          EL(zx, n, t, i) = zxv[i];
          EL(lx, n, t, i) = lxv[i];
          if (mkx) mkx[t * n + i] = (int8_t)mk;
        }
      }
      cs += state_cost<M>(d, zs, t, x);
      M::step(x, u, xn, d.dt);
#pragma unroll
      for (int i = 0; i < n; i++) x[i] = xn[i];
    }
    cost = cs + d.u_std * cc;
    if (d.n_obst > 0) {
      int i1, i2;
      lqt_obst_project<M>(d, c, &i1, &i2);
      if (c.valid && d.out.inner_iters) {
        d.out.inner_iters[((size_t)c.b * d.max_admm + it)] = i1 * 1000 + i2;      // set-convex iterations, Dykstra sweeps
      }
      const size_t arr = (size_t)d.T * d.N * n * TILE;
      const double *zn_a = c.at(d.obw + arr, d, n);
      for (int t = 0; t < d.N; t++) {
#pragma unroll
        for (int i = 0; i < n; i++) {                      // admm.py:49-59 with z = project_x(.)
          const double xv = EL(xa, n, t, i), zn = EL(zn_a, n, t, i), zo = EL(zx, n, t, i);
          const double r = __dsub_rn(xv, zn), dz = __dsub_rn(zn, zo);
          EL(lx, n, t, i) = __dadd_rn(EL(lx, n, t, i), r);
          EL(zx, n, t, i) = zn;
          prx = fma(r, r, prx);
          drx = fma(dz, dz, drx);
        }
      }
    }
    const double pprim = prim, pdual = dual;
    prim = sqrt(prx) + sqrt(pru);
    dual = sqrt(drx) + sqrt(dru);
    if (c.valid && d.out.res_log) {
      double *r = d.out.res_log + ((size_t)c.b * d.max_admm + it) * 2;
      r[0] = prim;
      r[1] = dual;
    }
    if (!d.fixed_budget) {
      if (prim < d.tol && dual < d.tol) ex = ISLS_ADMM_CONVERGED;
      else {
        const double pch = fabs(pprim - prim) / (pprim + 1e-30);
        const double dch = fabs(pdual - dual) / (pdual + 1e-30);
        if (pch < d.tol && dch < d.tol) ex = ISLS_ADMM_STALLED;
      }
    }
  }
  if (!ex) ex = ISLS_ADMM_MAXIT;
  d.cost[c.b] = cost;
  d.oit[c.b] = 1;
  d.nlog[c.b] = 2;
  d.odone[c.b] = 1;
  d.status[c.b] = 0;
  if (c.valid) {
    if (d.out.admm_iters) d.out.admm_iters[c.b] = it;
    if (d.out.admm_exit) d.out.admm_exit[c.b] = ex;
    d.out.cost_log[c.b * 2 + 0] = nan("");
    d.out.cost_log[c.b * 2 + 1] = cost;
  }
}

// broadcast lane 0 of tile 0 of the gain arrays into natural-layout K for every problem (LQT: shared gains)
template <class M>
__global__ void k_lqt_unpack_K(Dev d) {
  constexpr int n = M::n, m = M::m;
  constexpr int nt = NTRI(M::m);
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  const isls_solve_out &o = d.out;
  for (int t = 0; t < d.N; t++) {
    const bool last = t == d.N - 1;
    for (int i = 0; i < m * n; i++) {
      if (o.K) o.K[((size_t)b * d.N + t) * m * n + i] = EL(d.Kg, m * n, t, i);
      if (o.Qux) o.Qux[((size_t)b * d.N + t) * m * n + i] = last ? 0.0 : EL(d.Qux, m * n, t, i);
    }
    for (int a = 0; a < m; a++)
      for (int b2 = 0; b2 < m; b2++) {
        const int q = a >= b2 ? tri(a, b2) : tri(b2, a);
        if (o.Quu) o.Quu[(((size_t)b * d.N + t) * m + a) * m + b2] = last ? 0.0 : EL(d.Quu, nt, t, q);
        if (o.Quu_inv) o.Quu_inv[(((size_t)b * d.N + t) * m + a) * m + b2] = last ? 0.0 : EL(d.Qui, nt, t, q);
      }
  }
}

// ----------------------------------------------------------------------------------------------------- host side
struct isls_plan {
  isls_problem_desc desc;   // pointers inside are NOT valid after create (copied to the device block)
  int n, m, N, n_via, L, NJA;
  bool proj_x, proj_u;
  void *cblock;             // device constant block
  double lti[54];           // ISLS_MODEL_LTI: A [n, n] then B [n, m] at offset 36 (host copy, uploaded per launch sequence)
  Dev base;                 // constants filled in
};

// workspace carving (isls_b200.cu): returns total bytes; if base != NULL fills the Dev pointers
size_t isls_carve(const isls_plan *p, long long B, char *base, Dev *d, Dev *alt = nullptr);

// Opt-in for more than 48 KB of dynamic shared memory.  The attribute is per DEVICE (and per kernel), so it is tracked
// per (kernel instantiation, device id) - a process-wide flag left the second GPU of a process without it.
template <auto Kern>
static inline int ensure_dyn_smem(int bytes) {
  static unsigned long long done = 0;          // bit i: set on device i (benign race: setting twice is harmless)
  int dev = 0;
  CK(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  (void)bytes;                                 // opt in to the architectural maximum once: later launches of the same
  if (!(done & bit)) {                         // kernel may ask for more (the plan-constant block grows with N)
    cudaFuncAttributes fa;                     // static + dynamic shared memory share the 227 KB of a CTA
    CK(cudaFuncGetAttributes(&fa, Kern));
    CK(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
    done |= bit;
  }
  return 0;
}

// NaN / -1 fill of the optional ADMM logs (LQT path: written only up to each problem's exit iteration)
static __global__ void k_fill_logs(double *res_log, int *alpha_idx, long long cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  if (res_log) { res_log[2 * i] = nan(""); res_log[2 * i + 1] = nan(""); }
  if (alpha_idx) alpha_idx[i] = -1;
}

#define TPB_TILES 2   // tiles (warps) per CTA for the one-thread-per-problem kernels
static inline int n_tiles(const Dev &d) { return (d.tile1 - d.tile0 + d.tstep - 1) / d.tstep; }   // tiles of this launch
static dim3 tp_block() { return dim3(TILE, TPB_TILES); }
static dim3 tp_grid(const Dev &d) { return dim3((n_tiles(d) + TPB_TILES - 1) / TPB_TILES); }

static inline int ovl_env(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

// k_ff_tma launcher for one ring shape: launches iff the ring fits an SM and (unless forced) all tiles are resident in
// one wave.  Returns non-zero on a CUDA error; `launched` reports whether the kernel went out.
template <class M, bool PX, bool JC, int TC, int NST>
static int try_ff_tma(const Dev &d, cudaStream_t s, int tiles, int sms, bool force, bool &launched) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m), NJ = M::NJA;
  constexpr int SB = (JC ? NJ : 0) + ((!JC || PX) ? n : 0) + m + m * n + 2 * nt + m + (PX ? n : 0);
  constexpr size_t ring = (size_t)NST * TC * SB * TILE * sizeof(double);
  if constexpr (ring <= 200 * 1024) {
    const size_t smem = 128 + ring + (size_t)d.N * (m + (PX ? n : 0)) * sizeof(double) + 2 * (size_t)d.N * sizeof(int);
    if (smem > 226 * 1024) return 0;
    const long long per_sm = std::min<long long>(32, (227 * 1024) / (long long)(smem + 1024));
    if (!force && per_sm * sms < tiles) return 0;
    if (ensure_dyn_smem<k_ff_tma<M, PX, JC, TC, NST>>((int)smem)) return 1;
    k_ff_tma<M, PX, JC, TC, NST><<<tiles, TILE, smem, s>>>(d);
    launched = true;
  }
  return 0;
}

// K-pass launcher: below 1,536 tiles the latency-oriented form (plan constants one step ahead, Jacobian cache)
template <class M>
static void launch_kpass(const Dev &d, cudaStream_t s) {
  if (d.isls_C > 0) {           // isls_admm: closed-loop (Joseph) form of the value recursion, see riccati_step
    if (d.Jc) k_kpass<M, true, true><<<tp_grid(d), tp_block(), 0, s>>>(d);
    else k_kpass<M, false, true><<<tp_grid(d), tp_block(), 0, s>>>(d);
  } else if (d.Jc) k_kpass<M, true><<<tp_grid(d), tp_block(), 0, s>>>(d);
  else k_kpass<M, false><<<tp_grid(d), tp_block(), 0, s>>>(d);
}

// k_ff_ws launcher for one ring shape (see try_ff_tma)
template <class M, bool PX, bool JC, int G, int TC, int NST>
static int try_ff_ws(const Dev &d, cudaStream_t s, int tiles, int sms, bool force, bool &launched) {
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m), NJ = M::NJA;
  constexpr int SB = (JC ? NJ : 0) + ((!JC || PX) ? n : 0) + m + m * n + 2 * nt + m + (PX ? n : 0);
  constexpr size_t ring = (size_t)NST * TC * SB * TILE * sizeof(double);
  if constexpr (ring <= 200 * 1024) {
    const size_t smem = 128 + ring + 2 * n * TILE * sizeof(double) + (size_t)d.N * (m + (PX ? n : 0)) * sizeof(double) +
                        2 * (size_t)d.N * sizeof(int);
    if (smem > 226 * 1024) return 0;
    const long long per_sm = std::min<long long>(4, (227 * 1024) / (long long)(smem + 1024));
    if (!force && per_sm * sms < tiles) return 0;
    if (ensure_dyn_smem<k_ff_ws<M, PX, JC, G, TC, NST>>((int)smem)) return 1;
    k_ff_ws<M, PX, JC, G, TC, NST><<<tiles, dim3(TILE, G), smem, s>>>(d);
    launched = true;
  }
  return 0;
}

// ff-pass launcher: plain kernel for large batches (bandwidth-bound), cp.async-staged kernel for small ones
template <class M>
static int launch_ff(const Dev &d, cudaStream_t s) {
  static int mode = -2;
  if (mode == -2) {
    const char *e = getenv("ISLS_FF_STAGES");         // -1 auto (default), 0 plain, 2/3/4 forced pipeline depth
    mode = e ? atoi(e) : -1;
  }
  constexpr int n = M::n, m = M::m, nt = NTRI(M::m);
  constexpr int SB = n + m + m * n + 2 * nt + m + n, SF = m * n + 3 * m + n, SL = SB > SF ? SB : SF;
  const int tiles = n_tiles(d);
  // small batches (less than one warp per SM scheduler): TMA-staged single-warp CTAs, the deepest operand ring at
  // which all tiles are resident in one wave; Jacobian scalars from the cache of k_kpass when the workspace has one
  static int ff_mode = -2, ff_jc = -1;
  if (ff_mode == -2) {
    const char *e = getenv("ISLS_FF_MODE");           // -1 auto (default), 0: cp.async / plain rules below, 2: TMA
    ff_mode = e ? atoi(e) : -1;
    const char *j = getenv("ISLS_FF_JC");             // 0: recompute the Jacobian in the TMA kernels
    ff_jc = j ? atoi(j) : 1;
  }
  static const int ff_big = ovl_env("ISLS_FF_BIG", 22);   // >= 1,536 tiles: 0 plain k_ff, 22 / 13 / 12: k_ff_tma<TC, NST>
  int dev_ = 0, sms_ = 148;
  cudaGetDevice(&dev_);
  cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev_);
  // Mid-size batches (512 < tiles <= one wave of the shallow-ring form: 8 single-warp CTAs per SM at N = 100) are bandwidth-bound
  // already: the shallow-ring form without the Jacobian cache moves fewer bytes than the deep-ring form with it (car,
  // 32,768 problems: 178 vs 209 us per launch; at 1,280 tiles - two waves - it loses, 291 vs 253 us; at 512 tiles they tie)
  static const int ff_mid = ovl_env("ISLS_FF_MID", 1);
  const size_t smem22 = d.proj_x ? FfTmaShape<M, true, false, 2, 2>::smem_bytes(d.N)
                                 : FfTmaShape<M, false, false, 2, 2>::smem_bytes(d.N);
  const long long wave22 = (long long)sms_ * std::min<long long>(32, (227 * 1024) / (long long)(smem22 + 1024));
  const bool mid = ff_mid && M::n < 6 && tiles > 512 && tiles <= wave22;
  if (mode < 0 && ff_mode < 0 && (tiles >= isls_small_tiles() || mid) && ff_big > 0) {
    // large batches: the TMA-staged kernel also beats the plain one when HBM-bound (65,536 car problems: 0.356 vs
    // 0.389 ms; shallow rings = more resident single-warp CTAs per SM; profiles/r2_tuning_log.md), Jacobian recomputed
    const int sms = sms_;
    bool done = false;
    auto big = [&](auto px) -> int {
      constexpr bool PX_ = decltype(px)::value;
      if (ff_big == 13) return try_ff_tma<M, PX_, false, 1, 3>(d, s, tiles, sms, true, done);
      if (ff_big == 12) return try_ff_tma<M, PX_, false, 1, 2>(d, s, tiles, sms, true, done);
      return try_ff_tma<M, PX_, false, 2, 2>(d, s, tiles, sms, true, done);
    };
    if (d.proj_x ? big(std::true_type{}) : big(std::false_type{})) return 1;
    if (done) return 0;
  }
  if (mode < 0 && (ff_mode == 2 || (ff_mode < 0 && tiles < isls_small_tiles()))) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool jc = d.Jc != nullptr && ff_jc != 0, force = ff_mode == 2;
    bool done = false;
    if constexpr (M::n >= 6) {
      // larger models (arm): the rows of the n-sized algebra split over 4 warps per tile (k_ff_ws).  Off by default:
      // bit-identical but not faster - at C3's size the arm's ff-pass is HBM-bound (1.45 GB per launch at 5.9 TB/s),
      // at smaller ones its step time is the TMA round trip of the shallow ring its 17.7 KB per step leave room for
      // (profiles/r2_tuning_log.md, section 2)
      static const int ff_ws = ovl_env("ISLS_FF_WS", 0);
      if (ff_ws && jc && d.cost_kind == ISLS_COST_QUADRATIC) {
        auto gw = [&](auto px) -> int {
          constexpr bool PX_ = decltype(px)::value;
          if (try_ff_ws<M, PX_, true, 4, 2, 2>(d, s, tiles, sms, false, done) || done) return 0;
          if (try_ff_ws<M, PX_, true, 4, 1, 2>(d, s, tiles, sms, false, done) || done) return 0;
          return 0;
        };
        if (d.proj_x) gw(std::true_type{}); else gw(std::false_type{});
        if (done) return 0;
      }
    }
    auto go = [&](auto px, auto jcc) -> int {
      constexpr bool PX_ = decltype(px)::value, JC_ = decltype(jcc)::value;
      if (try_ff_tma<M, PX_, JC_, 4, 4>(d, s, tiles, sms, false, done) || done) return 0;
      if (try_ff_tma<M, PX_, JC_, 3, 3>(d, s, tiles, sms, false, done) || done) return 0;
      if (try_ff_tma<M, PX_, JC_, 2, 2>(d, s, tiles, sms, false, done) || done) return 0;
      if (try_ff_tma<M, PX_, JC_, 1, 2>(d, s, tiles, sms, force, done) || done) return 0;   // large slabs (arm: 17.7 KB per step)
      return 0;
    };
    if (d.proj_x) { if (jc) go(std::true_type{}, std::true_type{}); else go(std::true_type{}, std::false_type{}); }
    else { if (jc) go(std::false_type{}, std::true_type{}); else go(std::false_type{}, std::false_type{}); }
    if (done) return 0;
  }
  int stages = mode;
  // small batches are latency-bound: stage up to 4 steps ahead through shared memory (car 17 KB, arm 64.5 KB per
  // single-warp CTA at depth 4).  The depth is the deepest one at which ALL tiles are resident at once: with the arm's
  // 64.5 KB stage only 3 CTAs fit an SM (444 slots), so C3's 512 tiles ran as two waves (the second 15 % full, the
  // kernel twice one CTA's latency); depth 3 (48 KB, 4 CTAs per SM, 592 slots) runs them as one.
  if (stages < 0) {
    stages = 0;
    if (tiles < 1536) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const size_t per_sm = 227 * 1024, per_stage = (size_t)SL * TILE * sizeof(double);
      for (int st = 4; st >= 2 && !stages; st--) {
        const size_t cta = st * per_stage + 1024;                       // + the per-CTA reservation
        if (st == 4 && st * per_stage > 66 * 1024) continue;
        const long long slots = (long long)sms * (long long)std::min((size_t)32, per_sm / cta);
        if (st == 2 || slots >= tiles) stages = st;
      }
      if (stages == 2 && (size_t)2 * per_stage > 66 * 1024) stages = 0;
    }
  }
  if (stages == 4) {
    const size_t smem = (size_t)4 * SL * TILE * sizeof(double);
    if (ensure_dyn_smem<k_ff_staged<M, 4>>((int)smem)) return 1;
    k_ff_staged<M, 4><<<tiles, TILE, smem, s>>>(d);
  } else if (stages == 3) {
    const size_t smem = (size_t)3 * SL * TILE * sizeof(double);
    if (ensure_dyn_smem<k_ff_staged<M, 3>>((int)smem)) return 1;
    k_ff_staged<M, 3><<<tiles, TILE, smem, s>>>(d);
  } else if (stages == 2) {
    const size_t smem = (size_t)2 * SL * TILE * sizeof(double);
    if (ensure_dyn_smem<k_ff_staged<M, 2>>((int)smem)) return 1;
    k_ff_staged<M, 2><<<tiles, TILE, smem, s>>>(d);
  } else {
    k_ff<M><<<tp_grid(d), tp_block(), 0, s>>>(d);
  }
  return 0;
}


// Line-search CTA shape: CPT candidates per thread (independent FP64 chains, shared loads), W = ceil(L/CPT)
// warps.  MAXW only feeds __launch_bounds__ (register budget).
struct LsFuse { int fuse, outer, inner; };
template <class M, int CPT, int MAXW, int MINB = 1>
static void launch_ls_cfg(const Dev &d, bool closed, cudaStream_t s, LsFuse f) {
  const int W = (d.L + CPT - 1) / CPT;
  if (closed) k_linesearch_closed<M, CPT, MAXW, MINB><<<n_tiles(d), dim3(TILE, W), 0, s>>>(d);
  else if (d.proj_x) k_linesearch<M, CPT, MAXW, MINB, true><<<n_tiles(d), dim3(TILE, W), 0, s>>>(d, f.fuse, f.outer, f.inner);
  else k_linesearch<M, CPT, MAXW, MINB, false><<<n_tiles(d), dim3(TILE, W), 0, s>>>(d, f.fuse, f.outer, f.inner);
}
static int solve_compact() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("ISLS_COMPACT"); v = e ? atoi(e) : 1; }   // 0 off, k: every k-th outer iteration
  return v;
}
static int no_fused_update() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("ISLS_NO_FUSED_UPDATE"); v = e ? atoi(e) : 0; }
  return v;
}
static int ls_cpt_override() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("ISLS_LS_CPT");      // tuning knob: candidates per thread of the line-search CTA
    v = e ? atoi(e) : 0;
  }
  return v;
}
template <class M>
static void launch_linesearch(const Dev &d, bool closed, cudaStream_t s, LsFuse f = LsFuse{0, 0, 0}) {
  const int ov = ls_cpt_override();
  if constexpr (M::n >= 9) {
    if (d.L <= 8) launch_ls_cfg<M, 1, 8>(d, closed, s, f);
    else launch_ls_cfg<M, 2, 25>(d, closed, s, f);
  } else {
    if (d.L <= 20) {
      if (ov == 5) launch_ls_cfg<M, 5, 4, 2>(d, closed, s, f);
      else if (ov == 4) launch_ls_cfg<M, 4, 5, 3>(d, closed, s, f);
      else if (ov == 2) launch_ls_cfg<M, 2, 10, 2>(d, closed, s, f);
      else if (ov == 1) launch_ls_cfg<M, 1, 20, 1>(d, closed, s, f);
      else launch_ls_cfg<M, 5, 4, 3>(d, closed, s, f);      // 5 chains/thread, 4 warps, 3 CTAs/SM (168 regs)
    } else launch_ls_cfg<M, 5, 10>(d, closed, s, f);        // 21..50 candidates: 10 warps x 5 chains, one CTA per SM
  }
}

// ADMM warm start of the LQT path from natural-layout arrays (ADMM_LQT_Batch: the unconstrained solution, sls.py:266-268)
template <class M>
__global__ void k_pack_zinit(Dev d, const double *zx_in, const double *zu_in) {
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  double *zx = c.at(d.zx, d, M::n), *zu = c.at(d.zu, d, M::m);
  for (int t = 0; t < d.N; t++) {
    if (zx_in) for (int i = 0; i < M::n; i++) EL(zx, M::n, t, i) = zx_in[(c.ob * d.N + t) * M::n + i];
    if (zu_in) for (int j = 0; j < M::m; j++) EL(zu, M::m, t, j) = zu_in[(c.ob * d.N + t) * M::m + j];
  }
}

template <class M>
__global__ void k_pack_zs(Dev d, const double *zs_in) {
  const int tile = d.tile0 + (blockIdx.x * blockDim.y + threadIdx.y) * d.tstep;
  if (tile >= d.tile1) return;
  TileCtx<M> c(d, tile, threadIdx.x);
  double *zs = d.zs + (size_t)tile * d.n_via * M::n * TILE + c.lane;
  for (int k = 0; k < d.n_via; k++)
    for (int i = 0; i < M::n; i++) EL(zs, M::n, k, i) = zs_in[(c.ob * d.n_via + k) * M::n + i];
}

// ------------------------------------------------------------------------------------- per-model launch sequences
// One table of host entry points per dynamics model; isls_b200.cu looks the table up by model id.  Dev is passed by
// value (already set up: constants, workspace pointers, options, outputs).
struct isls_model_ops {
  int (*ilqr_admm)(const isls_plan *plan, Dev d, int64_t B, const double *x0, const double *u_init, const double *zs,
                   void *ws, cudaStream_t s);
  int (*ilqr)(Dev d, const double *x0, const double *u_init, const double *zs, cudaStream_t s);
  int (*isls_admm)(Dev d, SocSet S, SocX X, int64_t B, const double *x0, const double *u_init, const double *zs,
                   double *du_dev, double *phi_u_dev, cudaStream_t s);
  int (*lqt_admm)(Dev d, const isls_solve_opts *opts, int64_t B, const double *x0, const double *zs, cudaStream_t s);
  int (*rollout_linesearch)(Dev d, const double *x_nom, const double *u_nom, const double *du, const double *zs,
                            const double *reg_x, const double *reg_u, double *costs, int32_t *best, double *x_best,
                            double *u_best, cudaStream_t s);
  int (*overlap_probe)(Dev d, const double *x0, const double *u_init, const double *zs, cudaStream_t s, int ls_ctas,
                       int ff_depth, double *ms);
  int (*prepare)(const isls_plan *plan, cudaStream_t s);   // NULL, or stream-ordered set-up before a launch sequence (LTI)
};
const isls_model_ops *isls_ops_car();
const isls_model_ops *isls_ops_arm3();
const isls_model_ops *isls_ops_tassa_car();
const isls_model_ops *isls_ops_double_integrator(int m);
const isls_model_ops *isls_ops_lti(int n, int m);

// ---- overlapped schedule (large batches, control-only projections): the FP64-bound line search of one half of the
// tiles runs concurrently with the HBM-bound kernels (feed-forward pass; outer end + K-pass at outer boundaries) of the
// other half, on two streams.  The line search runs as a persistent kernel with a capped number of CTAs per SM so the
// TMA-staged feed-forward CTAs of the other half find registers and shared memory on every SM.
struct OvlAux {
  cudaStream_t s;
  cudaEvent_t ev[8];
};
static inline OvlAux *ovl_aux() {                        // per host thread and device: side stream + event ring
  static thread_local OvlAux aux[64];
  static thread_local bool have[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  dev &= 63;
  if (!have[dev]) {
    if (cudaStreamCreateWithFlags(&aux[dev].s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto &e : aux[dev].ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    have[dev] = true;
  }
  return &aux[dev];
}

template <class M>
struct ModelImpl {
  // iSLS.ilqr_admm (isls/isls.py:379-501): k_init, then per outer iteration K-pass, I_a x (ff-pass + linear rollout,
  // line search [+ fused ADMM update | winner rollout + ADMM update]), outer end; k_finalize.
  static int ilqr_admm(const isls_plan *plan, Dev d, int64_t B, const double *x0, const double *u_init,
                       const double *zs, void *ws, cudaStream_t s) {
    // Reference stop rules: finished problems are retired and the active ones re-packed into dense tiles after
    // every outer iteration (k_compact_*), so a tile never carries idle lanes for long.  (Cutting the batch into
    // chunks on separate streams was measured and dropped: every kernel fills the GPU, profiles/r1_tuning_log.md.)
    const bool compact = !d.fixed_budget && d.max_outer > 1 && solve_compact();
    Dev dalt = d;
    if (compact) isls_carve(plan, B, (char *)ws, &d, &dalt);
    else d.orig = nullptr;
    if constexpr (M::n < 9) {
      // experimental, off by default: measured slower than the sequential schedule (the line search needs the whole
      // register file for its throughput, so the two kernels of a slot do not co-run; profiles/r2_tuning_log.md)
      static const int ovl = ovl_env("ISLS_OVERLAP", 0);       // 0 off, 1 forced
      const bool fuse_ok = !d.proj_x && d.proj_u && !no_fused_update();
      if (ovl == 1 && !compact && !g_prof_on && fuse_ok && d.L <= 20 && d.T >= 2)
        return ilqr_admm_overlapped(d, x0, u_init, zs, s);
    }
    {
      Dev dc = d;
      cudaStream_t cs = s;
      LAUNCH(ISLS_KC_INIT, cs, (k_init<M><<<tp_grid(dc), tp_block(), 0, cs>>>(dc, x0, u_init, zs)));
      for (int j = 0; j < d.max_outer; j++) {
        LAUNCH(ISLS_KC_KPASS, cs, launch_kpass<M>(dc, cs));
        bool looped = false;
        if constexpr (M::n < 9) {
          // experimental, off by default (measured slower, profiles/r2_tuning_log.md section 5): the whole inner loop of
          // an outer iteration in one launch of persistent CTAs (k_admm_loop)
          static const int loop_mode = ovl_env("ISLS_ADMM_LOOP", 0);        // 0 off, 1 batches >= 1,536 tiles, 2 any size
          static const int loop_split = ovl_env("ISLS_ADMM_LOOP_SPLIT", 0); // 1: one launch per ADMM iteration
          static const int loop_stag = ovl_env("ISLS_ADMM_LOOP_STAGGER", 0);   // ns per arrival slot (first wave)
          static const int loop_grid = ovl_env("ISLS_ADMM_LOOP_GRID", 0);   // 0 auto (resident CTAs), else CTAs
          const bool ok = !d.proj_x && d.proj_u && !no_fused_update() && d.L <= 20 && d.cost_kind == ISLS_COST_QUADRATIC &&
                          (loop_mode == 2 || (loop_mode == 1 && n_tiles(dc) >= 1536));
          if (ok) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int tiles = n_tiles(dc);
            auto go = [&](auto jc, auto tc, auto nst) -> int {
              constexpr bool JC_ = decltype(jc)::value;
              constexpr int TC_ = decltype(tc)::value, NST_ = decltype(nst)::value;
              using FS = FfTmaShape<M, false, JC_, TC_, NST_>;
              constexpr auto kern = k_admm_loop<M, 5, 4, 3, JC_, TC_, NST_>;
              const int W = (d.L + 4) / 5;
              // W private operand rings + one copy of the plan constants (+ 23 KB of static line-search staging)
              const size_t smem = (size_t)W * (128 + FS::RING_BYTES) + (FS::smem_bytes(d.N) - 128 - FS::RING_BYTES);
              const long long per_sm = std::min<long long>(3, (227 * 1024) / (long long)(smem + 24 * 1024));
              if (per_sm < 1) return 0;
              const int grid = (int)std::min<long long>(tiles, loop_grid > 0 ? loop_grid : per_sm * sms);   // persistent
              if (ensure_dyn_smem<kern>((int)smem)) return 1;
              int *sm_ctr = dc.nact + 64;
              if (loop_stag) CK(cudaMemsetAsync(sm_ctr, 0, 256 * sizeof(int), cs));
              const dim3 blk(TILE, W);
              if (loop_split) {
                for (int a = 0; a < d.max_admm; a++)
                  LAUNCH(ISLS_KC_ADMM_LOOP, cs, (kern<<<grid, blk, smem, cs>>>(dc, j, a, a + 1, (unsigned)loop_stag, sm_ctr)));
              } else {
                LAUNCH(ISLS_KC_ADMM_LOOP, cs, (kern<<<grid, blk, smem, cs>>>(dc, j, 0, d.max_admm, (unsigned)loop_stag, sm_ctr)));
              }
              looped = true;
              return 0;
            };
            using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
            // small batches carry the Jacobian cache of k_kpass<., SMALL>
            if (dc.Jc ? go(std::true_type{}, I2{}, I2{}) : go(std::false_type{}, I1{}, I2{})) return 1;
          }
        }
        for (int a = 0; a < d.max_admm && !looped; a++) {
          const int fuse = (!d.proj_x && d.proj_u && !no_fused_update()) ? 1 : 0;      // streaming ADMM epilogue
          LAUNCH(ISLS_KC_FF, cs, launch_ff<M>(dc, cs));
          LAUNCH(ISLS_KC_LINESEARCH, cs, launch_linesearch<M>(dc, false, cs, LsFuse{fuse, j, a}));
          if (!fuse) {
            ProfScope ps__(ISLS_KC_ADMM, cs);
            launch_admm<M>(dc, j, a, cs);
            if (d.n_obst > 0)        // obstacle sets: all rows of a problem are projected together, CTA = problem
              k_obst_project<M><<<(unsigned)((dc.tile1 - dc.tile0) * TILE), ((d.N + 31) / 32) * 32, 0, cs>>>(dc, j, a);
          }
        }
        LAUNCH(ISLS_KC_OUTER_END, cs, (k_outer_end<M><<<tp_grid(dc), tp_block(), 0, cs>>>(dc, j)));
        if (compact && j + 1 < d.max_outer && (j + 1) % solve_compact() == 0) {
          // (measured on C5 with the reference stop rules: every iteration 103.1 ms, every 2nd 103.5, every 3rd 104.9,
          //  never 115.7 ms per 65,536 solves)
          ProfScope ps__(ISLS_KC_COMPACT, cs);
          isls_launch_compact_scan(dc.T * TILE, dc.odone, dc.orig, dc.newpos, dc.nact, cs);
          Dev db = dc;                      // alternate buffers become current
          db.xh = dalt.xh; db.uh = dalt.uh; db.zx = dalt.zx; db.zu = dalt.zu; db.zs = dalt.zs; db.cost = dalt.cost;
          db.nlog = dalt.nlog; db.status = dalt.status; db.oit = dalt.oit; db.orig = dalt.orig; db.odone = dalt.odone;
          k_compact_move<M><<<tp_grid(dc), tp_block(), 0, cs>>>(dc, db);
          dalt.xh = dc.xh; dalt.uh = dc.uh; dalt.zx = dc.zx; dalt.zu = dc.zu; dalt.zs = dc.zs; dalt.cost = dc.cost;
          dalt.nlog = dc.nlog; dalt.status = dc.status; dalt.oit = dc.oit; dalt.orig = dc.orig; dalt.odone = dc.odone;
          dc = db;
        }
      }
      LAUNCH(ISLS_KC_FINALIZE, cs, (k_finalize<M><<<tp_grid(dc), tp_block(), 0, cs>>>(dc)));
    }
    CK(cudaGetLastError());
    return ISLS_OK;
  }

  // Overlapped form of the same launch sequence.  The tiles are split into two interleaved halves H0 (even tiles), H1
  // (odd tiles) - two independent chains  K-pass, [ff, line search] x I_a, outer end, K-pass, ...  - and H1 runs one slot
  // behind H0, so in every slot one half is in its FP64-bound step (line search, stream A) while the other is in its
  // HBM-bound step (ff-pass; at outer boundaries outer end + K-pass + ff-pass; stream B).  A slot ends with an event
  // barrier between the two streams.  Same kernels and arithmetic as the sequential schedule: results are bit-identical.
  static int ilqr_admm_overlapped(Dev d, const double *x0, const double *u_init, const double *zs, cudaStream_t sA) {
    OvlAux *aux = ovl_aux();
    if (!aux) return fail(ISLS_E_INVALID, "could not create the side stream of the overlapped schedule");
    cudaStream_t sB = aux->s;
    static const int ls_ctas = ovl_env("ISLS_OVL_LS_CTAS", 2);       // resident line-search CTAs per SM
    static const int ff_depth = ovl_env("ISLS_OVL_FF_DEPTH", 4);     // feed-forward ring depth (steps in flight + 1)
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int evi = 0;
    auto barrier = [&]() -> int {                              // both streams wait for each other
      cudaEvent_t ea = aux->ev[evi & 7], eb = aux->ev[(evi + 1) & 7];
      evi += 2;
      CK(cudaEventRecord(ea, sA));
      CK(cudaEventRecord(eb, sB));
      CK(cudaStreamWaitEvent(sA, eb, 0));
      CK(cudaStreamWaitEvent(sB, ea, 0));
      return 0;
    };
    Dev h[2] = {d, d};
    for (int q = 0; q < 2; q++) { h[q].tile0 = d.tile0 + q; h[q].tstep = 2; }
    int *ctr = d.nact + 8;                                     // tile counter + exit count of the persistent line search
    CK(cudaMemsetAsync(d.nact, 0, 64 * sizeof(int), sA));
    k_init<M><<<tp_grid(d), tp_block(), 0, sA>>>(d, x0, u_init, zs);
    if (barrier()) return 1;
    const int I_a = d.max_admm, steps = 2 * d.max_outer * I_a + 1;   // per half: (HBM, LS) x I_o I_a, final outer end
    auto hbm_step = [&](const Dev &dh, int i) -> int {         // stream B: [outer end j-1,] [K-pass j,] ff-pass (j, a)
      const int j = i / I_a, a = i % I_a;
      if (a == 0) {
        if (j > 0) k_outer_end<M><<<tp_grid(dh), tp_block(), 0, sB>>>(dh, j - 1);
        launch_kpass<M>(dh, sB);
      }
      bool done = false;
      const int nt_ = n_tiles(dh);
      int rc = 0;
      if (ff_depth >= 6) rc = try_ff_tma<M, false, false, 1, 6>(dh, sB, nt_, sms, true, done);
      else if (ff_depth == 5) rc = try_ff_tma<M, false, false, 1, 5>(dh, sB, nt_, sms, true, done);
      else if (ff_depth == 4) rc = try_ff_tma<M, false, false, 1, 4>(dh, sB, nt_, sms, true, done);
      else if (ff_depth == 3) rc = try_ff_tma<M, false, false, 1, 3>(dh, sB, nt_, sms, true, done);
      else if (ff_depth > 0) rc = try_ff_tma<M, false, false, 2, 2>(dh, sB, nt_, sms, true, done);
      if (rc) return rc;
      if (!done) k_ff<M><<<tp_grid(dh), tp_block(), 0, sB>>>(dh);
      return 0;
    };
    auto ls_step = [&](const Dev &dh, int i) {                 // stream A: line search + fused ADMM update (j, a)
      const int j = i / I_a, a = i % I_a;
      const int grid = std::min(n_tiles(dh), ls_ctas * sms);
      if (ls_ctas > 0)
        k_linesearch_pers<M, 5, 4, 3, false><<<grid, dim3(TILE, (dh.L + 4) / 5), 0, sA>>>(dh, 1, j, a, ctr);
      else
        k_linesearch<M, 5, 4, 3, false><<<n_tiles(dh), dim3(TILE, (dh.L + 4) / 5), 0, sA>>>(dh, 1, j, a);
    };
    for (int t = 0; t <= steps; t++) {                         // slot t: H0 does its step t, H1 its step t - 1
      for (int q = 0; q < 2; q++) {
        const int k = t - q;
        if (k < 0 || k >= steps) continue;
        if (k == steps - 1) k_outer_end<M><<<tp_grid(h[q]), tp_block(), 0, sB>>>(h[q], d.max_outer - 1);
        else if (k & 1) ls_step(h[q], k / 2);
        else if (hbm_step(h[q], k / 2)) return 1;
      }
      if (barrier()) return 1;
    }
    k_finalize<M><<<tp_grid(d), tp_block(), 0, sA>>>(d);
    CK(cudaGetLastError());
    return ISLS_OK;
  }

  // iSLS.solve(method='dp') (isls/isls.py:54-132)
  static int ilqr(Dev d, const double *x0, const double *u_init, const double *zs, cudaStream_t s) {
    LAUNCH(ISLS_KC_INIT, s, (k_init<M><<<tp_grid(d), tp_block(), 0, s>>>(d, x0, u_init, zs)));
    for (int j = 0; j < d.max_outer; j++) {
      LAUNCH(ISLS_KC_BACKWARD_FULL, s, (k_backward_full<M><<<tp_grid(d), tp_block(), 0, s>>>(d)));
      LAUNCH(ISLS_KC_LINESEARCH, s, launch_linesearch<M>(d, true, s));
      LAUNCH(ISLS_KC_ACCEPT, s, (k_accept_closed<M><<<tp_grid(d), tp_block(), 0, s>>>(d, j)));
    }
    LAUNCH(ISLS_KC_FINALIZE, s, (k_finalize<M><<<tp_grid(d), tp_block(), 0, s>>>(d)));
    CK(cudaGetLastError());
    return ISLS_OK;
  }

  // iSLS.isls_admm (isls/isls.py:503-712)
  static int isls_admm(Dev d, SocSet S, SocX X, int64_t B, const double *x0, const double *u_init, const double *zs,
                       double *du_dev, double *phi_u_dev, cudaStream_t s) {
    const int rows = d.N * M::m, threads = ((rows + 31) / 32) * 32;
    LAUNCH(ISLS_KC_INIT, s, (k_init<M><<<tp_grid(d), tp_block(), 0, s>>>(d, x0, u_init, zs)));
    for (int j = 0; j < d.max_outer; j++) {
      LAUNCH(ISLS_KC_KPASS, s, launch_kpass<M>(d, s));
      k_isls_reset<M><<<tp_grid(d), tp_block(), 0, s>>>(d);
      for (int a = 0; a < d.max_admm; a++) {
        LAUNCH(ISLS_KC_FF, s, launch_ff<M>(d, s));
        LAUNCH(ISLS_KC_LINESEARCH, s, launch_linesearch<M>(d, false, s, LsFuse{0, j, a}));
        LAUNCH(ISLS_KC_ISLS_COLS, s, launch_isls_cols<M>(d, s));
        {
          ProfScope ps__(ISLS_KC_ISLS_UPDATE, s);
          if (threads <= 384 && S.P == 2 && S.c == 4 && S.ra == 5) k_isls_update<M, 2, 4, 5><<<(unsigned)B, threads, 0, s>>>(d, S, X, j, a);   // dim = 3
          else if (threads <= 384 && S.P == 2 && S.c == 3 && S.ra == 4) k_isls_update<M, 2, 3, 4><<<(unsigned)B, threads, 0, s>>>(d, S, X, j, a);
          else k_isls_update<M><<<(unsigned)B, threads, 0, s>>>(d, S, X, j, a);
        }
      }
      LAUNCH(ISLS_KC_OUTER_END, s, (k_outer_end<M><<<tp_grid(d), tp_block(), 0, s>>>(d, j)));
    }
    LAUNCH(ISLS_KC_FINALIZE, s, (k_finalize<M><<<tp_grid(d), tp_block(), 0, s>>>(d)));
    k_isls_out<M><<<tp_grid(d), tp_block(), 0, s>>>(d, du_dev, phi_u_dev);
    CK(cudaGetLastError());
    return ISLS_OK;
  }

  // SLS.ADMM_LQT_DP / ADMM_LQT_Batch (isls/sls.py:250-317)
  static int lqt_admm(Dev d, const isls_solve_opts *opts, int64_t B, const double *x0, const double *zs,
                      cudaStream_t s) {
    const size_t T = d.T, N = d.N;
    CK(cudaMemsetAsync(d.zx, 0, T * N * M::n * TILE * sizeof(double), s));
    CK(cudaMemsetAsync(d.lx, 0, T * N * M::n * TILE * sizeof(double), s));
    CK(cudaMemsetAsync(d.zu, 0, T * N * M::m * TILE * sizeof(double), s));
    CK(cudaMemsetAsync(d.lu, 0, T * N * M::m * TILE * sizeof(double), s));
    CK(cudaMemsetAsync(d.xh, 0, N * M::n * TILE * sizeof(double), s));      // tile 0: linearisation point (unused)
    CK(cudaMemsetAsync(d.uh, 0, N * M::m * TILE * sizeof(double), s));
    CK(cudaMemsetAsync(d.odone, 0, T * TILE * sizeof(int), s));
    CK(cudaMemsetAsync(d.status, 0, T * TILE * sizeof(int), s));
    CK(cudaMemsetAsync(d.cost, 0, T * TILE * sizeof(double), s));
    if (d.out.inner_iters) CK(cudaMemsetAsync(d.out.inner_iters, 0, (size_t)B * d.max_admm * sizeof(int), s));
    // the ADMM logs are written up to each problem's exit iteration only: NaN / -1 beyond it, like the iLQR paths
    if (d.out.res_log || d.out.alpha_idx)
      k_fill_logs<<<(unsigned)((B * d.max_admm + 255) / 256), 256, 0, s>>>(d.out.res_log, d.out.alpha_idx,
                                                                           (long long)B * d.max_admm);
    // shared gains: K-pass on tile 0 only (linear model: the Jacobian does not depend on the trajectory)
    Dev d1 = d;
    d1.T = 1;
    d1.tile1 = 1;
    d1.B = 1;
    k_kpass<M><<<1, dim3(TILE, 1), 0, s>>>(d1);
    // the K-pass reset (lambda = 0, reg = z) touched tile 0 only with zeros: state stays zero
    k_pack_zs<M><<<tp_grid(d), tp_block(), 0, s>>>(d, zs);
    if (opts->z_x_init_dev || opts->z_u_init_dev)
      k_pack_zinit<M><<<tp_grid(d), tp_block(), 0, s>>>(d, opts->z_x_init_dev, opts->z_u_init_dev);
    {
      constexpr int nt_ = NTRI(M::m);
      const size_t lqt_smem = N * (size_t)(2 * M::m * M::n + 2 * nt_ + 4 * M::n + 3 * M::m) * sizeof(double) +
                              2 * N * sizeof(int);
      static int lqt_mode = -2;
      if (lqt_mode == -2) {
        const char *e = getenv("ISLS_LQT_SMEM");          // -1 auto (default), 0 global-memory constants
        lqt_mode = e ? atoi(e) : -1;
      }
      if (lqt_mode != 0 && lqt_smem <= 200 * 1024) {
        if (ensure_dyn_smem<k_lqt_admm<M, true>>(200 * 1024)) return 1;
        LAUNCH(ISLS_KC_LQT, s, (k_lqt_admm<M, true><<<tp_grid(d), tp_block(), lqt_smem, s>>>(d, x0)));
      } else {
        LAUNCH(ISLS_KC_LQT, s, (k_lqt_admm<M, false><<<tp_grid(d), tp_block(), 0, s>>>(d, x0)));
      }
    }
    Dev df = d;
    df.out.K = nullptr;          // gains and Riccati logs are shared: unpacked by k_lqt_unpack_K
    df.out.Qux = df.out.Quu = df.out.Quu_inv = nullptr;
    k_finalize<M><<<tp_grid(d), tp_block(), 0, s>>>(df);
    if (d.out.K || d.out.Qux || d.out.Quu || d.out.Quu_inv)
      k_lqt_unpack_K<M><<<(unsigned)((B + 127) / 128), 128, 0, s>>>(d);
    CK(cudaGetLastError());
    return ISLS_OK;
  }

  // stage level: iSLS.rollout_batch + cost + argmin (isls/isls.py:135-154, 468-477)
  static int rollout_linesearch(Dev d, const double *x_nom, const double *u_nom, const double *du, const double *zs,
                                const double *reg_x, const double *reg_u, double *costs, int32_t *best,
                                double *x_best, double *u_best, cudaStream_t s) {
    k_pack_stage<M><<<tp_grid(d), tp_block(), 0, s>>>(d, x_nom, u_nom, du, zs, reg_x, reg_u);
    launch_linesearch<M>(d, false, s);
    k_unpack_stage<M><<<tp_grid(d), tp_block(), 0, s>>>(d, costs, best, x_best, u_best);
    CK(cudaGetLastError());
    return ISLS_OK;
  }

  // Measurement helper (isls_probe_overlap_f64): durations of the two kernels of one slot of the overlapped schedule, each
  // alone on its half of the tiles and both together on two streams.  ms[0] line search (one CTA per tile), ms[1]
  // persistent line search with ls_ctas CTAs per SM, ms[2] TMA-staged ff-pass (ring depth ff_depth), ms[3] plain ff-pass,
  // ms[4] persistent line search || TMA ff-pass, ms[5] plain line search || TMA ff-pass.  Best of 3, synchronous.
  static int overlap_probe(Dev d, const double *x0, const double *u_init, const double *zs, cudaStream_t sA, int ls_ctas,
                           int ff_depth, double *ms) {
    if constexpr (M::n >= 9) {
      return fail(ISLS_E_UNSUPPORTED, "overlap probe: car-sized models only");
    } else {
      OvlAux *aux = ovl_aux();
      if (!aux) return fail(ISLS_E_INVALID, "could not create the side stream");
      cudaStream_t sB = aux->s;
      int dev = 0, sms = 148;
      CK(cudaGetDevice(&dev));
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      d.orig = nullptr;
      Dev h[2] = {d, d};
      for (int q = 0; q < 2; q++) { h[q].tile0 = d.tile0 + q; h[q].tstep = 2; }
      int *ctr = d.nact + 8;
      CK(cudaMemsetAsync(d.nact, 0, 64 * sizeof(int), sA));
      k_init<M><<<tp_grid(d), tp_block(), 0, sA>>>(d, x0, u_init, zs);
      launch_kpass<M>(d, sA);
      k_ff<M><<<tp_grid(d), tp_block(), 0, sA>>>(d);
      const dim3 lsb(TILE, (d.L + 4) / 5);
      auto ls_plain = [&](cudaStream_t s) { k_linesearch<M, 5, 4, 3, false><<<n_tiles(h[0]), lsb, 0, s>>>(h[0], 1, 0, 0); };
      auto ls_pers = [&](cudaStream_t s) {
        k_linesearch_pers<M, 5, 4, 3, false><<<std::min(n_tiles(h[0]), std::max(1, ls_ctas) * sms), lsb, 0, s>>>(h[0], 1, 0, 0, ctr);
      };
      auto ff_tma = [&](cudaStream_t s) -> int {
        bool done = false;
        const int nt_ = n_tiles(h[1]);
        if (ff_depth >= 6) return try_ff_tma<M, false, false, 1, 6>(h[1], s, nt_, sms, true, done);
        if (ff_depth == 5) return try_ff_tma<M, false, false, 1, 5>(h[1], s, nt_, sms, true, done);
        if (ff_depth == 4) return try_ff_tma<M, false, false, 1, 4>(h[1], s, nt_, sms, true, done);
        if (ff_depth == 3) return try_ff_tma<M, false, false, 1, 3>(h[1], s, nt_, sms, true, done);
        return try_ff_tma<M, false, false, 2, 2>(h[1], s, nt_, sms, true, done);
      };
      auto ff_plain = [&](cudaStream_t s) { k_ff<M><<<tp_grid(h[1]), tp_block(), 0, s>>>(h[1]); };
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      auto timed = [&](int idx, auto &&fa, auto &&fb) -> int {      // fa on stream A, fb (may be empty) on stream B
        double best = 1e30;
        for (int rep = 0; rep < 4; rep++) {
          CK(cudaStreamSynchronize(sA));
          CK(cudaStreamSynchronize(sB));
          CK(cudaEventRecord(e0, sA));
          CK(cudaStreamWaitEvent(sB, e0, 0));
          fa(sA);
          fb(sB);
          CK(cudaEventRecord(aux->ev[0], sB));
          CK(cudaStreamWaitEvent(sA, aux->ev[0], 0));
          CK(cudaEventRecord(e1, sA));
          CK(cudaEventSynchronize(e1));
          float t = 0.f;
          CK(cudaEventElapsedTime(&t, e0, e1));
          if (rep > 0 && t < best) best = t;
        }
        ms[idx] = best;
        return 0;
      };
      auto none = [&](cudaStream_t) {};
      auto fft = [&](cudaStream_t s) { ff_tma(s); };
      if (timed(0, ls_plain, none)) return 1;
      if (timed(1, ls_pers, none)) return 1;
      if (timed(2, none, fft)) return 1;
      if (timed(3, none, ff_plain)) return 1;
      if (timed(4, ls_pers, fft)) return 1;
      if (timed(5, ls_plain, fft)) return 1;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      CK(cudaGetLastError());
      return ISLS_OK;
    }
  }

  static const isls_model_ops *ops() {
    static const isls_model_ops o = {&ilqr_admm, &ilqr, &isls_admm, &lqt_admm, &rollout_linesearch, &overlap_probe, nullptr};
    return &o;
  }
};
