// Second-order-cone row projections shared by ADMM_SLS (sls.cu) and the robust iSLS-ADMM (isls_b200.cu):
// project_set_convex (isls/projections.py:289-374) with projections = [project_soc_unit] * P, one row per thread, one
// problem per CTA (the reference's stop rule is a maximum over ALL rows of the problem, SURVEY D9 batch semantics).
#pragma once
#include <math.h>
#include <string.h>

#include <algorithm>

#define SOC_MAXC 4      // columns of the projected rows: 1 + x_dim/2
#define SOC_MAXR 5      // rows of each A_i (= c + 1)
#define SOC_MAXP 4      // cones per row
struct SocSet {
  int P, c, ra;                       // number of cones, columns, rows of A_i
  double A[SOC_MAXP][SOC_MAXR][SOC_MAXC];
  double b[SOC_MAXP][SOC_MAXR];
  double linv[SOC_MAXC][SOC_MAXC];    // (I + rho sum A_i'A_i)^-1
  double rho, threshold;
  int max_iter;
};


// State side of the robust iSLS-ADMM (isls.py:631-638): the rows of the listed state components are projected, one
// project_set_convex call per component over its N rows, onto the cones of S with the component's own offsets.
#define SOC_MAXCOMP 8
struct SocX {
  int ncomp;                 // 0: no state projection
  int u_identity;            // 1: the control side is not projected (isls_admm without project_u): z_u = y_u
  int comp[SOC_MAXCOMP];
  double b[SOC_MAXCOMP][SOC_MAXP][SOC_MAXR];
};

__device__ __forceinline__ double block_max(double v, double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = sm[0];
  for (int i = 1; i < (int)((blockDim.x + 31) >> 5); i++) r = fmax(r, sm[i]);
  return r;
}
__device__ __forceinline__ double block_sum(double v, double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < (int)((blockDim.x + 31) >> 5); i++) r += sm[i];
  return r;
}

// project_soc_unit_batch (isls/projections.py:140-162) on one row [z(0..d-1), t], incl. the D9 behaviour (every
// row with t < 0 is zeroed): later masks win (cond2, then cond1, then cond3), exactly like the numpy code.
__device__ __forceinline__ void soc_unit_row(int d, const double *y, double *out) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < SOC_MAXR; i++) if (i < d) s = fma(y[i], y[i], s);
  const double zn = sqrt(s), t = y[d];
  const bool c1 = (zn <= -t) || (t < 0.0);
  const bool c2 = (zn > t) || (zn > -t);
  const bool c3 = zn <= t;
  for (int i = 0; i <= d; i++) out[i] = y[i];
  if (c2) {
    const double tmp = (zn + t) / 2.0;
    for (int i = 0; i < d; i++) out[i] = tmp * y[i] / (zn + 1e-30);
    out[d] = tmp;
  }
  if (c1) for (int i = 0; i <= d; i++) out[i] = 0.0;
  if (c3) for (int i = 0; i <= d; i++) out[i] = y[i];
}


// Inner ADMM of project_set_convex for this thread's row x0[c] -> x[c]; `act` = the thread owns a row; every thread of
// the CTA must call it (block reductions).  Returns the number of inner iterations.
// `bb`: the offsets b_i to use (S.b, or a row's own); red == nullptr: a single row projected on its own (its stop rule
// is its own residual; no block reduction, callable by a subset of the CTA).
// CP, CC, CR: compile-time number of cones, columns and rows of A_i (0 = take them from S at run time).  With the
// compile-time form every loop unrolls and z_i, lambda_i, x live in registers; the run-time form indexes them
// dynamically, i.e. in local memory - 0.9 GB of DRAM writes per ADMM_SLS call at C4 (profiles/r2_ncu_sls_kernels.csv).
template <int CP = 0, int CC = 0, int CR = 0>
__device__ __forceinline__ int soc_project_set(const SocSet &S, const double (&bb)[SOC_MAXP][SOC_MAXR],
                                               const double (&x0)[SOC_MAXC], double (&x)[SOC_MAXC], bool act,
                                               double *red) {
  const int c = CC ? CC : S.c, SP = CP ? CP : S.P, Sra = CR ? CR : S.ra;
  double zi[SOC_MAXP][SOC_MAXR], li[SOC_MAXP][SOC_MAXR];
  int inner = 0;
  _Pragma("unroll") for (int q = 0; q < c; q++) x[q] = x0[q];
  _Pragma("unroll") for (int i = 0; i < SP; i++)
    _Pragma("unroll") for (int e = 0; e < Sra; e++) {
      double v = bb[i][e];
      _Pragma("unroll") for (int q = 0; q < c; q++) v = fma(S.A[i][e][q], x[q], v);
      zi[i][e] = v;                         // z_i = A_i x + b_i   (projections.py:315)
      li[i][e] = 0.0;
    }
  double pm = 1e5, dm = 1e5;
  for (int j = 0; j < S.max_iter; j++) {
    inner++;
    double rsd[SOC_MAXC] = {};
    _Pragma("unroll") for (int i = 0; i < SP; i++)
      _Pragma("unroll") for (int e = 0; e < Sra; e++) {
        const double w = (-bb[i][e] + zi[i][e]) - li[i][e];
        _Pragma("unroll") for (int q = 0; q < c; q++) rsd[q] = fma(S.A[i][e][q], w, rsd[q]);
      }
    double tq[SOC_MAXC];
    _Pragma("unroll") for (int q = 0; q < c; q++) tq[q] = x0[q] + S.rho * rsd[q];
    _Pragma("unroll") for (int q = 0; q < c; q++) {
      double v = 0.0;
      _Pragma("unroll") for (int p = 0; p < c; p++) v = fma(S.linv[q][p], tq[p], v);
      x[q] = v;                              // projections.py:330
    }
    double pmax = 0.0, dmax = 0.0;
    _Pragma("unroll") for (int i = 0; i < SP; i++) {
      double axb[SOC_MAXR], y[SOC_MAXR], zn[SOC_MAXR];
      _Pragma("unroll") for (int e = 0; e < Sra; e++) {
        double v = bb[i][e];
        _Pragma("unroll") for (int q = 0; q < c; q++) v = fma(S.A[i][e][q], x[q], v);
        axb[e] = v;
        y[e] = v + li[i][e];
      }
      soc_unit_row(Sra - 1, y, zn);
      double ps = 0.0, dr[SOC_MAXC] = {};
      _Pragma("unroll") for (int e = 0; e < Sra; e++) {
        const double pr = axb[e] - zn[e];
        ps = fma(pr, pr, ps);
        const double dz = zn[e] - zi[i][e];
        _Pragma("unroll") for (int q = 0; q < c; q++) dr[q] = fma(S.A[i][e][q], dz, dr[q]);
        li[i][e] += pr;
        zi[i][e] = zn[e];
      }
      double ds = 0.0;
      _Pragma("unroll") for (int q = 0; q < c; q++) ds = fma(S.rho * dr[q], S.rho * dr[q], ds);
      pmax = fmax(pmax, sqrt(ps));
      dmax = fmax(dmax, sqrt(ds));
    }
    // stop rule on the max over rows and cones (projections.py:343-348)
    const double pprev = pm, dprev = dm;
    pm = red ? block_max(act ? pmax : 0.0, red) : pmax;
    dm = red ? block_max(act ? dmax : 0.0, red) : dmax;
    if (pm < S.threshold && dm < S.threshold) break;
    if (j < S.max_iter - 1) {
      const double pch = fabs(pprev - pm) / (pprev + 1e-30), dch = fabs(dprev - dm) / (dprev + 1e-30);
      if (pch < 1e-5 && dch < 1e-5) break;
    }
  }
  return inner;
}

// (I + rho sum A_i'A_i)^-1 and the cone data of a SocSet from host arrays As [P, ra, c], bs [P, ra]
static inline void soc_set_build(SocSet *Sp, int P, int c, int ra, const double *As, const double *bs, double rho,
                                 int max_iter, double threshold) {
  SocSet &S = *Sp;
  memset(&S, 0, sizeof(S));
  S.P = P; S.c = c; S.ra = ra; S.rho = rho; S.threshold = threshold; S.max_iter = max_iter;
  // (I + rho sum A_i'A_i)^-1 on the host (c x c, c <= 4): Gauss-Jordan with partial pivoting
  double Mx[SOC_MAXC][2 * SOC_MAXC] = {};
  for (int i = 0; i < S.P; i++)
    for (int e = 0; e < S.ra; e++) {
      S.b[i][e] = bs[i * S.ra + e];
      for (int q = 0; q < c; q++) S.A[i][e][q] = As[(i * S.ra + e) * c + q];
    }
  for (int q = 0; q < c; q++)
    for (int r2 = 0; r2 < c; r2++) {
      double v = (q == r2) ? 1.0 : 0.0;
      for (int i = 0; i < S.P; i++)
        for (int e = 0; e < S.ra; e++) v += S.rho * S.A[i][e][q] * S.A[i][e][r2];
      Mx[q][r2] = v;
      Mx[q][c + r2] = (q == r2) ? 1.0 : 0.0;
    }
  for (int col = 0; col < c; col++) {
    int piv = col;
    for (int r2 = col + 1; r2 < c; r2++) if (fabs(Mx[r2][col]) > fabs(Mx[piv][col])) piv = r2;
    for (int k = 0; k < 2 * c; k++) std::swap(Mx[col][k], Mx[piv][k]);
    const double d = Mx[col][col];
    for (int k = 0; k < 2 * c; k++) Mx[col][k] /= d;
    for (int r2 = 0; r2 < c; r2++)
      if (r2 != col) {
        const double f = Mx[r2][col];
        for (int k = 0; k < 2 * c; k++) Mx[r2][k] -= f * Mx[col][k];
      }
  }
  for (int q = 0; q < c; q++)
    for (int r2 = 0; r2 < c; r2++) S.linv[q][r2] = Mx[q][c + r2];
}
