// SLS path of libisls_b200.so (item 4 of the north star, BASELINE config 4): SLS.solve_sls, SLS.ADMM_SLS with
// second-order-cone chance constraints, SLS.controller  (isls/sls.py:205-242, 319-454; isls/base.py:32-50, 98-119;
// isls/projections.py:140-162, 289-374).
//
// B200-first re-design, not a port of the numpy code:
//   * the reference inverts every trailing principal sub-matrix L[i m:, i m:] by N successive Woodbury down-dates
//     (base.py:32-50) and multiplies each with one block column of -Su'Q Sw (sls.py:226-229).  Here ONE "reverse"
//     Cholesky factorisation L = U U' (U upper triangular) serves all trailing sub-matrices at once, because
//     L[k:,k:] = U[k:,k:] U[k:,k:]'; with W = U^-1 the whole block-lower-triangular PHI_U is two masked dense
//     products  PHI_U = blt( W' * blt(W * r) )  - the dense contractions run on the FP64 tensor pipe (DMMA
//     mma.sync.m8n8k4.f64; tcgen05 has no f64 kind).
//   * ADMM_SLS: one CTA per problem, one thread per row of [d_u | Phi_u(:, :n/2)]; the (N m)^2 operator is
//     shared by the batch (L2 resident); the row-wise SOC projections (inner ADMM of project_set_convex, stop rule
//     = max over rows and cones) use block-wide reductions.
//   * controller: PHI_X = Sw + Su PHI_U is unit block lower triangular, so K = PHI_U PHI_X^-1 is a block
//     back-substitution per problem (the reference forms a dense LU inverse of (N n)^2).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/isls_b200.h"
#include "common.cuh"

// ---------------------------------------------------------------------------------------------- DMMA GEMM
// C[M x N] = alpha * op(A) * B (+ beta * C), row-major, op(A) = A [M x K] or A' with A stored [K x M].
// CTA = 4 warps computing a 32 x 32 tile, each warp a 16 x 16 sub-tile as 2 x 2 mma.m8n8k4.f64 fragments.
#define GT 32
#define GK 16
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) k_dgemm(int M, int N, int K, double alpha, const double *A, int lda, int transA,
                                               const double *B, int ldb, double beta, double *C, int ldc) {
  __shared__ double As[GT][GK + 1];
  __shared__ double Bs[GK][GT + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;
  double acc[2][2][2] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int e = tid; e < GT * GK; e += 128) {
      const int r = e / GK, c = e % GK;
      const int gm = m0 + r, gk = k0 + c;
      double v = 0.0;
      if (gm < M && gk < K) v = transA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[r][c] = v;
      const int br = e / GT, bc = e % GT;
      const int bk = k0 + br, bn = n0 + bc;
      Bs[br][bc] = (bk < K && bn < N) ? B[(size_t)bk * ldb + bn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; i++) a[i] = As[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 2; j++) b[j] = Bs[kk + (lane & 3)][wn + 8 * j + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int gm = m0 + wm + 8 * i + (lane >> 2), gn = n0 + wn + 8 * j + 2 * (lane & 3) + q;
        if (gm < M && gn < N) {
          double v = alpha * acc[i][j][q];
          if (beta != 0.0) v += beta * C[(size_t)gm * ldc + gn];
          C[(size_t)gm * ldc + gn] = v;
        }
      }
}

static void dgemm(cudaStream_t s, int M, int N, int K, double alpha, const double *A, int lda, bool transA,
                  const double *B, int ldb, double beta, double *C, int ldc) {
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT);
  k_dgemm<<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, transA ? 1 : 0, B, ldb, beta, C, ldc);
}

// ------------------------------------------------------------------------------------------- operator build
// Sw, Su of Base.AB (isls/base.py:98-119): Sw block (i,j) = A^(i-j) (i >= j), Su block (i,j) = A^(i-j-1) B (i > j).
// Apow[k] = A^k is built by repeated right-multiplication, the order of the reference's backward block recursion.
__global__ void k_sls_powers(int n, int N, const double *A, double *Apow) {
  // one CTA, n*n threads
  const int i = threadIdx.x / n, j = threadIdx.x % n;
  if (threadIdx.x < n * n) Apow[i * n + j] = (i == j) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = 1; k < N; k++) {
    if (threadIdx.x < n * n) {
      double acc = 0.0;
      for (int q = 0; q < n; q++) acc = fma(Apow[(size_t)(k - 1) * n * n + i * n + q], A[q * n + j], acc);
      Apow[(size_t)k * n * n + i * n + j] = acc;
    }
    __syncthreads();
  }
}

__global__ void k_sls_fill(int n, int m, int N, const double *Apow, const double *Bm, double *Sw, double *Su) {
  const int Nn = N * n, Nm = N * m;
  const size_t tot = (size_t)Nn * Nn;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / Nn), c = (int)(e % Nn);
    const int bi = r / n, bj = c / n;
    Sw[e] = (bi >= bj) ? Apow[(size_t)(bi - bj) * n * n + (r % n) * n + (c % n)] : 0.0;
  }
  const size_t totu = (size_t)Nn * Nm;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < totu; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / Nm), c = (int)(e % Nm);
    const int bi = r / n, bj = c / m;
    double v = 0.0;
    if (bi > bj) {
      const double *P = Apow + (size_t)(bi - bj - 1) * n * n + (r % n) * n;
      for (int q = 0; q < n; q++) v = fma(P[q], Bm[q * m + (c % m)], v);
    }
    Su[e] = v;
  }
}

// DTQ = Su' Q with diagonal Q (q[N n]):  DTQ[c][r] = Su[r][c] * q[r]
__global__ void k_sls_dtq(int Nn, int Nm, const double *Su, const double *q, double *DTQ) {
  const size_t tot = (size_t)Nn * Nm;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / Nn), r = (int)(e % Nn);
    DTQ[e] = Su[(size_t)r * Nm + c] * q[r];
  }
}

__global__ void k_add_diag(int n, double *Mx, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Mx[(size_t)i * n + i] += v;
}

// Reverse Cholesky L = U U' (U upper triangular) in place on a copy of L, then W = U^-1 (upper triangular).
// One CTA; the matrix lives in global memory (L2).  Sets *flag = 1 if a pivot is not positive.
__global__ void __launch_bounds__(1024) k_rchol_inv(int n, double *Amat, double *W, int *flag) {
  __shared__ double sd;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = n - 1; j >= 0; j--) {
    if (tid == 0) {
      double d = Amat[(size_t)j * n + j];
      if (!(d > 0.0)) { *flag = 1; d = 1.0; }
      sd = sqrt(d);
      Amat[(size_t)j * n + j] = sd;
    }
    __syncthreads();
    const double inv = 1.0 / sd;
    for (int i = tid; i < j; i += nt) Amat[(size_t)i * n + j] *= inv;     // U[i][j], i < j
    __syncthreads();
    // trailing (leading) update: A[i][k] -= U[i][j] U[k][j] for i <= k < j (upper part only)
    const int cnt = j * j;
    for (int e = tid; e < cnt; e += nt) {
      const int i = e / j, k = e % j;
      if (i <= k) Amat[(size_t)i * n + k] = fma(-Amat[(size_t)i * n + j], Amat[(size_t)k * n + j], Amat[(size_t)i * n + k]);
    }
    __syncthreads();
  }
  // W = U^-1: thread per column, back substitution
  for (int j = tid; j < n; j += nt) {
    for (int i = 0; i < n; i++) W[(size_t)i * n + j] = 0.0;
    W[(size_t)j * n + j] = 1.0 / Amat[(size_t)j * n + j];
    for (int i = j - 1; i >= 0; i--) {
      double acc = 0.0;
      for (int k = i + 1; k <= j; k++) acc = fma(Amat[(size_t)i * n + k], W[(size_t)k * n + j], acc);
      W[(size_t)i * n + j] = -acc / Amat[(size_t)i * n + i];
    }
  }
}

// zero the strictly-upper block part of a [N m x N n] matrix: rows of block-row bi < block-column bj
__global__ void k_mask_blt(int n, int m, int N, double *Y) {
  const size_t tot = (size_t)N * m * N * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / ((size_t)N * n)), c = (int)(e % ((size_t)N * n));
    if (r / m < c / n) Y[e] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------ ADMM_SLS
#include "soc.cuh"

struct SocRowsX {                       // state-side projection: a few rows of [d_x | Phi_x(:, :c-1)]
  int nx;
  int row[ISLS_MAX_XROWS];
  double qr[ISLS_MAX_XROWS];
  double b[ISLS_MAX_XROWS][SOC_MAXP][SOC_MAXR];
};

struct SlsAdmm {
  int Nm, Nn, c, max_iter, fixed_budget;
  double rho_u, alpha, tol;
  const double *linv;     // [Nm x Nm] (L + rho_u I)^-1, shared
  const double *DTQ;      // [Nm x Nn]
  const double *rfb;      // [Nm x (c-1)]  -Su'Q Sx, shared
  const double *xd;       // [B x Nn]
  double *du;             // [B x Nm]
  double *phic;           // [B x Nm x (c-1)]
  double *logs;           // [B x max_iter x 2] or NULL
  int *iters, *exits;     // [B]
  long long *inner_total; // [B] or NULL
  const double *Su;       // [Nn x Nm] (state side)
  const double *Sw;       // [Nn x Nn]
};

// One CTA per problem, thread r = row r of [d_u | Phi_u(:, :c-1)]   (sls.py:372-447 + projections.py:289-374)
// CP, CC, CR: compile-time cone-set shape (0: run time), see soc_project_set
template <int CP, int CC, int CR>
__global__ void k_sls_admm(SlsAdmm a, SocSet S, SocRowsX X) {
  extern __shared__ double sh[];
  double *rhs = sh;                         // [Nm][c] right-hand side of the current iteration, then x_u
  double *red = sh + (size_t)a.Nm * a.c;    // reduction scratch [32]
  double *zxs = red + 32;                   // [nx][c] z_x - lambda_x of the projected state rows (reg_x)
  double zx[SOC_MAXC] = {}, lx[SOC_MAXC] = {};           // thread i < nx owns state row X.row[i]
  const int b = blockIdx.x, r = threadIdx.x, c = CC ? CC : a.c;
  const bool act = r < a.Nm;
  // r_side row: column 0 = (Su'Q xd)[r], columns 1.. = -Su'Q Sx (shared)
  double rs[SOC_MAXC] = {}, z[SOC_MAXC] = {}, lm[SOC_MAXC] = {}, xu[SOC_MAXC] = {};
  if (act) {
    double acc = 0.0;
    const double *xd = a.xd + (size_t)b * a.Nn;
    for (int k = 0; k < a.Nn; k++) acc = fma(a.DTQ[(size_t)r * a.Nn + k], xd[k], acc);
    rs[0] = acc;
    _Pragma("unroll") for (int q = 1; q < c; q++) rs[q] = a.rfb[(size_t)r * (c - 1) + q - 1];
  }
  double prim = 1e6, dual = 1e6;
  int ex = 0, it = 0;
  long long inner = 0;
  for (it = 0; it < a.max_iter && !ex; it++) {
    // x_u = l_inv (r_side + Rr (z - lambda))                                   sls.py:372-380
    if (r < X.nx) _Pragma("unroll") for (int q = 0; q < c; q++) zxs[r * c + q] = zx[q] - lx[q];
    __syncthreads();
    if (act) _Pragma("unroll") for (int q = 0; q < c; q++) {
      double v = rs[q] + a.rho_u * (z[q] - lm[q]);
      for (int i = 0; i < X.nx; i++)                     // + Su'Qr reg_x (sls.py:371): Qr is zero off the listed rows
        v = fma(X.qr[i] * a.Su[(size_t)X.row[i] * a.Nm + r], zxs[i * c + q], v);
      rhs[r * c + q] = v;
    }
    __syncthreads();
    if (act) {
      double acc[SOC_MAXC] = {};
      const double *row = a.linv + (size_t)r * a.Nm;
      for (int k = 0; k < a.Nm; k++) {
        const double l = row[k];
        _Pragma("unroll") for (int q = 0; q < c; q++) acc[q] = fma(l, rhs[k * c + q], acc[q]);
      }
      _Pragma("unroll") for (int q = 0; q < c; q++) xu[q] = acc[q];
    }
    double psx = 0.0, dsx = 0.0;
    if (X.nx > 0) {
      // ---- state side: x_x = Su x_u (+ Sx on the feedback columns) at the listed rows, own projection each
      __syncthreads();
      if (act) _Pragma("unroll") for (int q = 0; q < c; q++) rhs[r * c + q] = xu[q];
      __syncthreads();
      int xin = 0;
      if (r < X.nx) {
        double xx[SOC_MAXC] = {}, y[SOC_MAXC], zn[SOC_MAXC];
        const double *su = a.Su + (size_t)X.row[r] * a.Nm;
        for (int k = 0; k < a.Nm; k++) {
          const double sv_ = su[k];
          _Pragma("unroll") for (int q = 0; q < c; q++) xx[q] = fma(sv_, rhs[k * c + q], xx[q]);
        }
        _Pragma("unroll") for (int q = 1; q < c; q++) xx[q] += a.Sw[(size_t)X.row[r] * a.Nn + q - 1];              // sls.py:381-382
        _Pragma("unroll") for (int q = 0; q < c; q++) y[q] = (a.alpha * xx[q] + (1.0 - a.alpha) * zx[q]) + lx[q];
        xin = soc_project_set<CP, CC, CR>(S, X.b[r], y, zn, true, nullptr);
        _Pragma("unroll") for (int q = 0; q < c; q++) {
          const double pr = xx[q] - zn[q], dz = zn[q] - zx[q];
          lx[q] += pr;
          zx[q] = zn[q];
          psx = fma(X.qr[r] * pr, X.qr[r] * pr, psx);
          dsx = fma(X.qr[r] * dz, X.qr[r] * dz, dsx);
        }
      }
      inner += (long long)(block_sum((double)xin, red) + 0.5);
    }
    // ---- z = project_set_convex(alpha x + (1-alpha) z + lambda)               sls.py:403-405
    double x0[SOC_MAXC], x[SOC_MAXC];
    _Pragma("unroll") for (int q = 0; q < c; q++) x0[q] = (a.alpha * xu[q] + (1.0 - a.alpha) * z[q]) + lm[q];
    inner += soc_project_set<CP, CC, CR>(S, S.b, x0, x, act, red);
    // ---- dual update and residuals (sls.py:406-418), Rr = rho_u I
    double ps = 0.0, ds = 0.0;
    _Pragma("unroll") for (int q = 0; q < c; q++) {
      const double zn = x[q], pr = xu[q] - zn, dz = zn - z[q];
      lm[q] += pr;
      z[q] = zn;
      ps = fma(a.rho_u * pr, a.rho_u * pr, ps);
      ds = fma(a.rho_u * dz, a.rho_u * dz, ds);
    }
    const double pprim = prim, pdual = dual;
    prim = sqrt(block_sum(act ? ps : 0.0, red));
    dual = sqrt(block_sum(act ? ds : 0.0, red));
    if (X.nx > 0) {                                                          // sls.py:414-415
      prim = sqrt(block_sum(psx, red)) + prim;
      dual = sqrt(block_sum(dsx, red)) + dual;
    }
    if (a.logs && r == 0) {
      a.logs[((size_t)b * a.max_iter + it) * 2] = prim;
      a.logs[((size_t)b * a.max_iter + it) * 2 + 1] = dual;
    }
    if (!a.fixed_budget) {
      if (prim < a.tol && dual < a.tol) ex = ISLS_ADMM_CONVERGED;
      else {
        const double pch = fabs(pprim - prim) / (pprim + 1e-30), dch = fabs(pdual - dual) / (pdual + 1e-30);
        if (pch < 1e-2 && dch < 1e-2) ex = ISLS_ADMM_STALLED;               // sls.py:429
      }
    }
  }
  if (!ex) ex = ISLS_ADMM_MAXIT;
  if (act) {
    a.du[(size_t)b * a.Nm + r] = xu[0];                                      // sls.py:449
    _Pragma("unroll") for (int q = 1; q < c; q++) a.phic[((size_t)b * a.Nm + r) * (c - 1) + q - 1] = xu[q];
  }
  if (r == 0) {
    a.iters[b] = it;
    a.exits[b] = ex;
    if (a.inner_total) a.inner_total[b] = inner;
  }
}

// du = Linv0 (Su'Q xd) for every problem  (sls.py:221): thread per (problem,row)
__global__ void k_sls_du(int Nm, int Nn, long long B, const double *Linv0, const double *DTQ, const double *xd,
                         double *tmp, double *du) {
  const long long b = blockIdx.x;
  extern __shared__ double sv[];       // [Nm]
  const double *x = xd + (size_t)b * Nn;
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < Nn; k++) acc = fma(DTQ[(size_t)r * Nn + k], x[k], acc);
    sv[r] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < Nm; k++) acc = fma(Linv0[(size_t)r * Nm + k], sv[k], acc);
    du[(size_t)b * Nm + r] = acc;
  }
}

// ---------------------------------------------------------------------------------------------- controller
// SLS.controller (sls.py:235-242) for one problem per CTA.  PHI_U = [phic (per problem, first c-1 columns) |
// PHI_U shared (remaining columns)].  PHI_X = Sw + Su PHI_U is unit block lower triangular, so K PHI_X = PHI_U is
// solved column block by column block from the right:  K[:, j] = PHI_U[:, j] - sum_{i > j} K[:, i] PHI_X[i, j].
// PHI_X columns are formed on the fly: PHI_X[:, j] = Sw[:, j] + Su PHI_U[:, j].
__global__ void k_sls_controller(int n, int m, int N, int cfirst, const double *Sw, const double *Su,
                                 const double *PHI_shared, const double *phic, const double *du, double *PHIX_ws,
                                 double *K, double *kff, size_t sw_stride = 0, size_t su_stride = 0) {
  const int Nn = N * n, Nm = N * m;
  const long long b = blockIdx.x;
  Sw += (size_t)b * sw_stride;          // per-problem operators (time-varying C, D of iSLS.controller); 0 = shared
  Su += (size_t)b * su_stride;
  double *PX = PHIX_ws + (size_t)b * Nn * Nn;
  double *Kb = K + (size_t)b * Nm * Nn;
  auto phi = [&](int r, int col) -> double {
    return col < cfirst ? phic[((size_t)b * Nm + r) * cfirst + col] : PHI_shared[(size_t)r * Nn + col];
  };
  // PHI_X = Sw + Su PHI_U   (dense per problem; (N n)^2 (N m) FMAs, L2-resident operands)
  for (int e = threadIdx.x; e < Nn * Nn; e += blockDim.x) {
    const int r = e / Nn, col = e % Nn;
    double acc = Sw[(size_t)r * Nn + col];
    // PHI_U[k, col] is zero for rows k above block column col; Su[r, k] is zero for k-blocks >= r-block
    const int k0 = (col / n) * m, k1 = (r / n) * m;
    for (int k = k0; k < k1; k++) acc = fma(Su[(size_t)r * Nm + k], phi(k, col), acc);
    PX[e] = acc;
  }
  __syncthreads();
  // back-substitution from the last column: thread per row of K
  for (int col = Nn - 1; col >= 0; col--) {
    for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
      double acc = phi(r, col);
      // K[r, i] nonzero only for i-blocks <= r-block (causal); PHI_X[i, col] nonzero for i >= col
      const int i1 = min(Nn, (r / m + 1) * n);
      for (int i = col + 1; i < i1; i++) acc = fma(-Kb[(size_t)r * Nn + i], PX[(size_t)i * Nn + col], acc);
      // diagonal entry of PHI_X: unit block lower triangular -> within the diagonal block PHI_X = I + 0
      Kb[(size_t)r * Nn + col] = (col < i1) ? acc : 0.0;
    }
    __syncthreads();
  }
  // k = (I - K Su) du = du - K (Su du)
  extern __shared__ double sv[];        // [Nn]
  const double *d = du + (size_t)b * Nm;
  for (int r = threadIdx.x; r < Nn; r += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < Nm; k++) acc = fma(Su[(size_t)r * Nm + k], d[k], acc);
    sv[r] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = d[r];
    for (int k = 0; k < Nn; k++) acc = fma(-Kb[(size_t)r * Nn + k], sv[k], acc);
    kff[(size_t)b * Nm + r] = acc;
  }
}

// C = (I - Z A_d)^-1 and D = C Z B_d of iSLSBase.AB (isls/isls_base.py:138-158) for time-varying A_t, B_t: block (r, j)
// of C is A_{r-1} A_{r-2} ... A_j (identity for r = j), block (r, j) of D is C(r, j+1) B_j (r > j), accumulated from
// the diagonal to the left exactly like the reference's loop (C[., j] = C[., j+1] @ A_j).  CTA = problem, thread = one
// row (r, i) of C and D, which it walks from its diagonal block leftwards with a row vector v = e_i' A_{r-1} ... A_j.
#define TV_MAXN 16
__global__ void k_tv_build_CD(int n, int m, int N, const double *A, const double *Bm, size_t a_stride, size_t b_stride,
                              double *C, double *D) {
  const int Nn = N * n, Nm = N * m;
  const long long b = blockIdx.x;
  A += (size_t)b * a_stride;
  Bm += (size_t)b * b_stride;
  double *Cb = C + (size_t)b * Nn * Nn, *Db = D + (size_t)b * Nn * Nm;
  for (int row = threadIdx.x; row < Nn; row += blockDim.x) {
    const int r = row / n, i = row % n;
    double v[TV_MAXN], w[TV_MAXN];
    for (int q = 0; q < n; q++) v[q] = (q == i) ? 1.0 : 0.0;
    for (int q = 0; q < Nn; q++) Cb[(size_t)row * Nn + q] = 0.0;
    for (int q = 0; q < Nm; q++) Db[(size_t)row * Nm + q] = 0.0;
    Cb[(size_t)row * Nn + row] = 1.0;
    for (int j = r - 1; j >= 0; j--) {
      const double *Aj = A + (size_t)j * n * n, *Bj = Bm + (size_t)j * n * m;
      for (int q = 0; q < m; q++) {
        double acc = 0.0;
        for (int k = 0; k < n; k++) acc = fma(v[k], Bj[k * m + q], acc);
        Db[(size_t)row * Nm + j * m + q] = acc;
      }
      for (int q = 0; q < n; q++) {
        double acc = 0.0;
        for (int k = 0; k < n; k++) acc = fma(v[k], Aj[k * n + q], acc);
        w[q] = acc;
      }
      for (int q = 0; q < n; q++) { v[q] = w[q]; Cb[(size_t)row * Nn + j * n + q] = w[q]; }
    }
  }
}

// Batched controller.  PHI_U of problem b differs from the shared PHI_U in its first `cfirst` columns only, so the
// columns j >= cfirst of PHI_X = Sw + Su PHI_U and - because the back-substitution runs from the right - of
// K = PHI_U PHI_X^-1 are THE SAME for every problem: k_sls_ctrl_shared computes them once per (plan, cfirst) with the
// arithmetic of k_sls_controller (same summation order, bit-identical), k_sls_ctrl_batch adds the cfirst per-problem
// columns (two (N n) x (N m) mat-vecs each instead of a dense (N n)^2 (N m) product and a 200-step substitution per
// problem: 4 MFLOP -> 0.1 MFLOP per problem at N = 50), assembles K[b] and k[b] = d_u - K (Su d_u).
__global__ void k_sls_ctrl_shared(int n, int m, int N, int cfirst, const double *Sw, const double *Su,
                                  const double *PHI_shared, double *PX, double *Ksh) {
  const int Nn = N * n, Nm = N * m;
  for (int e = threadIdx.x; e < Nn * Nn; e += blockDim.x) {
    const int r = e / Nn, col = e % Nn;
    double acc = Sw[(size_t)r * Nn + col];
    const int k0 = (col / n) * m, k1 = (r / n) * m;
    if (col >= cfirst)
      for (int k = k0; k < k1; k++) acc = fma(Su[(size_t)r * Nm + k], PHI_shared[(size_t)k * Nn + col], acc);
    PX[e] = acc;
  }
  __syncthreads();
  for (int col = Nn - 1; col >= cfirst; col--) {
    for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
      double acc = PHI_shared[(size_t)r * Nn + col];
      const int i1 = min(Nn, (r / m + 1) * n);
      for (int i = col + 1; i < i1; i++) acc = fma(-Ksh[(size_t)r * Nn + i], PX[(size_t)i * Nn + col], acc);
      Ksh[(size_t)r * Nn + col] = (col < i1) ? acc : 0.0;
    }
    __syncthreads();
  }
}

__global__ void k_sls_ctrl_batch(int n, int m, int N, int cfirst, const double *Sw, const double *Su, const double *Ksh,
                                 const double *phic, const double *du, double *K, double *kff) {
  const int Nn = N * n, Nm = N * m;
  const long long b = blockIdx.x;
  extern __shared__ double sh[];
  double *pxc = sh;                 // [Nn] current per-problem column of PHI_X; later Su d_u
  double *kown = sh + Nn;           // [cfirst][Nm] per-problem columns of K
  double *Kb = K + (size_t)b * Nm * Nn;
  for (int col = cfirst - 1; col >= 0; col--) {
    for (int i = threadIdx.x; i < Nn; i += blockDim.x) {
      double acc = Sw[(size_t)i * Nn + col];
      const int k0 = (col / n) * m, k1 = (i / n) * m;
      for (int k = k0; k < k1; k++) acc = fma(Su[(size_t)i * Nm + k], phic[((size_t)b * Nm + k) * cfirst + col], acc);
      pxc[i] = acc;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
      double acc = phic[((size_t)b * Nm + r) * cfirst + col];
      const int i1 = min(Nn, (r / m + 1) * n);
      for (int i = col + 1; i < i1; i++) {
        const double kri = i < cfirst ? kown[i * Nm + r] : Ksh[(size_t)r * Nn + i];
        acc = fma(-kri, pxc[i], acc);
      }
      kown[col * Nm + r] = (col < i1) ? acc : 0.0;
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < Nm * Nn; e += blockDim.x) {
    const int r = e / Nn, col = e % Nn;
    Kb[e] = col < cfirst ? kown[col * Nm + r] : Ksh[e];
  }
  const double *d = du + (size_t)b * Nm;
  for (int r = threadIdx.x; r < Nn; r += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < Nm; k++) acc = fma(Su[(size_t)r * Nm + k], d[k], acc);
    pxc[r] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = d[r];
    for (int k = 0; k < Nn; k++) {
      const double krk = k < cfirst ? kown[k * Nm + r] : Ksh[(size_t)r * Nn + k];
      acc = fma(-krk, pxc[k], acc);
    }
    kff[(size_t)b * Nm + r] = acc;
  }
}

// ----------------------------------------------------------------------------------------------------- host
struct isls_sls_plan {
  int n, m, N, Nn, Nm;
  double u_std;
  double *Apow, *Sw, *Su, *DTQ, *L, *Linv0, *W, *R, *Y, *PHI, *fac, *Lrho, *Linv_rho, *rfb, *q;
  double *PXsh, *Ksh;       // shared columns of PHI_X and K of the batched controller (valid for cfirst = ctrl_cfirst)
  int ctrl_cfirst;          // -1: not built yet
  int *flag;
  double rho_cached;
  int c_cached;
  bool have_rho;
};

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

// inverse of an SPD matrix through the reverse Cholesky factor: Minv = W' W with W = U^-1, M = U U'
static int spd_inverse_dev(isls_sls_plan *p, const double *Mx, double *Minv, cudaStream_t s) {
  const int n = p->Nm;
  CK(cudaMemcpyAsync(p->fac, Mx, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(p->flag, 0, sizeof(int), s));
  k_rchol_inv<<<1, 1024, 0, s>>>(n, p->fac, p->W, p->flag);
  dgemm(s, n, n, n, 1.0, p->W, n, true, p->W, n, 0.0, Minv, n);
  CK(cudaGetLastError());
  return ISLS_OK;
}

extern "C" int isls_sls_plan_create(int32_t n, int32_t m, int32_t N, const double *A_host, const double *B_host,
                                    const double *Qdiag_t_host, double u_std, isls_sls_plan **plan, void *stream) {
  if (!A_host || !B_host || !Qdiag_t_host || !plan || n < 1 || m < 1 || N < 2)
    return isls_fail(ISLS_E_INVALID, "NULL argument or bad size");
  if (n * n > 1024) return isls_fail(ISLS_E_UNSUPPORTED, "x_dim too large for the SLS operator build");
  cudaStream_t s = (cudaStream_t)stream;
  isls_sls_plan *p = new isls_sls_plan();
  memset(p, 0, sizeof(*p));
  p->n = n; p->m = m; p->N = N; p->Nn = N * n; p->Nm = N * m; p->u_std = u_std;
  const size_t Nn = p->Nn, Nm = p->Nm;
  double *dA, *dB;
  struct { double **ptr; size_t cnt; } arrs[] = {
      {&p->Apow, (size_t)N * n * n}, {&p->Sw, Nn * Nn}, {&p->Su, Nn * Nm}, {&p->DTQ, Nm * Nn}, {&p->L, Nm * Nm},
      {&p->Linv0, Nm * Nm}, {&p->W, Nm * Nm}, {&p->R, Nm * Nn}, {&p->Y, Nm * Nn}, {&p->PHI, Nm * Nn},
      {&p->fac, Nm * Nm}, {&p->Lrho, Nm * Nm}, {&p->Linv_rho, Nm * Nm}, {&p->rfb, Nm * (size_t)n}, {&p->q, Nn},
      {&p->PXsh, Nn * Nn}, {&p->Ksh, Nm * Nn}, {&dA, (size_t)n * n}, {&dB, (size_t)n * m}};
  p->ctrl_cfirst = -1;
  for (auto &a : arrs) CK(cudaMalloc(a.ptr, al(a.cnt * sizeof(double))));
  CK(cudaMalloc(&p->flag, 256));
  CK(cudaMemcpyAsync(dA, A_host, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dB, B_host, (size_t)n * m * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(p->q, Qdiag_t_host, Nn * sizeof(double), cudaMemcpyHostToDevice, s));
  k_sls_powers<<<1, std::max(32, n * n), 0, s>>>(n, N, dA, p->Apow);
  k_sls_fill<<<296, 256, 0, s>>>(n, m, N, p->Apow, dB, p->Sw, p->Su);
  k_sls_dtq<<<296, 256, 0, s>>>((int)Nn, (int)Nm, p->Su, p->q, p->DTQ);
  // L = Su'Q Su + R                                                          (sls.py:218)
  dgemm(s, (int)Nm, (int)Nm, (int)Nn, 1.0, p->DTQ, (int)Nn, false, p->Su, (int)Nm, 0.0, p->L, (int)Nm);
  k_add_diag<<<(unsigned)((Nm + 127) / 128), 128, 0, s>>>((int)Nm, p->L, u_std);
  int rc = spd_inverse_dev(p, p->L, p->Linv0, s);       // also leaves W = U^-1 of L
  if (rc) return rc;
  // r_side = -Su'Q Sw ; PHI_U = blt( W' * blt( W * r_side ) )               (sls.py:225-229 re-designed)
  dgemm(s, (int)Nm, (int)Nn, (int)Nn, -1.0, p->DTQ, (int)Nn, false, p->Sw, (int)Nn, 0.0, p->R, (int)Nn);
  dgemm(s, (int)Nm, (int)Nn, (int)Nm, 1.0, p->W, (int)Nm, false, p->R, (int)Nn, 0.0, p->Y, (int)Nn);
  k_mask_blt<<<296, 256, 0, s>>>(n, m, N, p->Y);
  dgemm(s, (int)Nm, (int)Nn, (int)Nm, 1.0, p->W, (int)Nm, true, p->Y, (int)Nn, 0.0, p->PHI, (int)Nn);
  k_mask_blt<<<296, 256, 0, s>>>(n, m, N, p->PHI);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  int flag = 0;
  CK(cudaMemcpy(&flag, p->flag, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(dA);
  cudaFree(dB);
  if (flag) { isls_sls_plan_destroy(p); return isls_fail(ISLS_E_INVALID, "Su'Q Su + R is not positive definite"); }
  *plan = p;
  return ISLS_OK;
}

extern "C" int isls_sls_plan_destroy(isls_sls_plan *p) {
  if (!p) return ISLS_OK;
  double *ptrs[] = {p->Apow, p->Sw, p->Su, p->DTQ, p->L, p->Linv0, p->W, p->R, p->Y, p->PHI, p->fac, p->Lrho,
                    p->Linv_rho, p->rfb, p->q, p->PXsh, p->Ksh};
  for (double *q : ptrs) cudaFree(q);
  cudaFree(p->flag);
  delete p;
  return ISLS_OK;
}

extern "C" int isls_sls_operators(const isls_sls_plan *p, double *Sw, double *Su, double *PHI_U, void *stream) {
  if (!p) return isls_fail(ISLS_E_INVALID, "plan is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t Nn = p->Nn, Nm = p->Nm;
  if (Sw) CK(cudaMemcpyAsync(Sw, p->Sw, Nn * Nn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (Su) CK(cudaMemcpyAsync(Su, p->Su, Nn * Nm * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (PHI_U) CK(cudaMemcpyAsync(PHI_U, p->PHI, Nm * Nn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return ISLS_OK;
}

extern "C" int isls_sls_solve_f64(const isls_sls_plan *p, int64_t B, const double *xd_dev, double *du_dev,
                                  void *stream) {
  if (!p || B <= 0 || !xd_dev || !du_dev) return isls_fail(ISLS_E_INVALID, "NULL argument or B <= 0");
  k_sls_du<<<(unsigned)B, 128, p->Nm * sizeof(double), (cudaStream_t)stream>>>(p->Nm, p->Nn, B, p->Linv0, p->DTQ,
                                                                               xd_dev, nullptr, du_dev);
  CK(cudaGetLastError());
  return ISLS_OK;
}

// state-side terms of the operators (sls.py:342-347): Lrho += sum_i qr_i Su[i,:]' Su[i,:],  rfb -= sum_i qr_i Su[i,:]' Sx[i,:]
__global__ void k_sls_xrows_update(int Nm, int Nn, int c1, SocRowsX X, const double *Su, const double *Sw, double *Lrho,
                                   double *rfb) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)Nm * Nm; e += (size_t)gridDim.x * blockDim.x) {
    const int ra = (int)(e / Nm), cb = (int)(e % Nm);
    double acc = 0.0;
    for (int i = 0; i < X.nx; i++) acc = fma(X.qr[i] * Su[(size_t)X.row[i] * Nm + ra], Su[(size_t)X.row[i] * Nm + cb], acc);
    Lrho[e] += acc;
    if (cb < c1) {
      double f = 0.0;
      for (int i = 0; i < X.nx; i++) f = fma(X.qr[i] * Su[(size_t)X.row[i] * Nm + ra], Sw[(size_t)X.row[i] * Nn + cb], f);
      rfb[(size_t)ra * c1 + cb] -= f;
    }
  }
}

extern "C" int isls_sls_admm_f64(isls_sls_plan *p, const isls_sls_admm_opts *o, int64_t B, const double *xd_dev,
                                 double *du_dev, double *phi_cols_dev, double *logs_dev, int32_t *iters_dev,
                                 int32_t *exit_dev, int64_t *inner_total_dev, void *stream) {
  if (!p || !o || B <= 0 || !xd_dev || !du_dev || !phi_cols_dev || !iters_dev || !exit_dev)
    return isls_fail(ISLS_E_INVALID, "NULL argument or B <= 0");
  if (o->struct_size != (uint32_t)sizeof(isls_sls_admm_opts))
    return isls_fail(ISLS_E_INVALID, "isls_sls_admm_opts.struct_size mismatch (binding built against another isls_b200.h?)");
  const int c = p->n / 2 + 1;
  if (o->n_cones < 1 || o->n_cones > SOC_MAXP || c > SOC_MAXC || o->cone_rows != c + 1 || !o->As || !o->bs)
    return isls_fail(ISLS_E_UNSUPPORTED, "unsupported cone set (need A_i of shape [c+1, c], c = 1 + x_dim/2 <= 4)");
  if (p->Nm > 1024) return isls_fail(ISLS_E_UNSUPPORTED, "N*u_dim > 1024");
  cudaStream_t s = (cudaStream_t)stream;
  SocRowsX X;
  memset(&X, 0, sizeof(X));
  if (o->n_x_rows < 0 || o->n_x_rows > ISLS_MAX_XROWS) return isls_fail(ISLS_E_INVALID, "n_x_rows out of range");
  if (o->n_x_rows > 0) {
    if (!o->x_row_idx || !o->x_bs || !o->rho_x_rows) return isls_fail(ISLS_E_INVALID, "state projection: NULL row data");
    if (o->n_x_rows > p->Nm) return isls_fail(ISLS_E_UNSUPPORTED, "more projected state rows than control rows");
    X.nx = o->n_x_rows;
    for (int i = 0; i < X.nx; i++) {
      if (o->x_row_idx[i] < 0 || o->x_row_idx[i] >= p->Nn) return isls_fail(ISLS_E_INVALID, "state row index out of range");
      X.row[i] = o->x_row_idx[i];
      X.qr[i] = o->rho_x_rows[i];
      for (int q = 0; q < o->n_cones; q++)
        for (int e = 0; e < o->cone_rows; e++) X.b[i][q][e] = o->x_bs[((size_t)i * o->n_cones + q) * o->cone_rows + e];
    }
  }
  // (L + rho_u I [+ Su'Qr Su])^-1 and -Su'Q Sx [- Su'Qr Sx], cached per (rho_u) when there is no state side
  if (!p->have_rho || p->rho_cached != o->rho_u || X.nx > 0) {
    CK(cudaMemcpyAsync(p->Lrho, p->L, (size_t)p->Nm * p->Nm * sizeof(double), cudaMemcpyDeviceToDevice, s));
    k_add_diag<<<(p->Nm + 127) / 128, 128, 0, s>>>(p->Nm, p->Lrho, o->rho_u);
    int rc = spd_inverse_dev(p, p->Lrho, p->Linv_rho, s);
    if (rc) return rc;
    // rfb = -DTQ * Sw[:, :c-1]  (Sx = first x_dim/2 columns of Sw)
    dgemm(s, p->Nm, c - 1, p->Nn, -1.0, p->DTQ, p->Nn, false, p->Sw, p->Nn, 0.0, p->rfb, c - 1);
    if (X.nx > 0) {
      CK(cudaMemcpyAsync(p->Lrho, p->L, (size_t)p->Nm * p->Nm * sizeof(double), cudaMemcpyDeviceToDevice, s));
      k_add_diag<<<(p->Nm + 127) / 128, 128, 0, s>>>(p->Nm, p->Lrho, o->rho_u);
      k_sls_xrows_update<<<148, 256, 0, s>>>(p->Nm, p->Nn, c - 1, X, p->Su, p->Sw, p->Lrho, p->rfb);
      int rc2 = spd_inverse_dev(p, p->Lrho, p->Linv_rho, s);
      if (rc2) return rc2;
    }
    // restore W = U^-1 of L (the shared PHI_U does not need it any more, but keep the plan consistent)
    p->rho_cached = o->rho_u;
    p->have_rho = X.nx == 0;
  }
  SocSet S;
  soc_set_build(&S, o->n_cones, c, o->cone_rows, o->As, o->bs, o->inner_rho, o->inner_max_iter, o->inner_threshold);
  SlsAdmm a;
  a.Nm = p->Nm; a.Nn = p->Nn; a.c = c; a.max_iter = o->max_iter; a.fixed_budget = o->fixed_budget;
  a.rho_u = o->rho_u; a.alpha = o->alpha; a.tol = o->tol;
  a.linv = p->Linv_rho; a.DTQ = p->DTQ; a.rfb = p->rfb; a.xd = xd_dev; a.du = du_dev; a.phic = phi_cols_dev;
  a.logs = logs_dev; a.iters = iters_dev; a.exits = exit_dev; a.inner_total = (long long *)inner_total_dev;
  a.Su = p->Su; a.Sw = p->Sw;
  const int threads = ((p->Nm + 31) / 32) * 32;
  const size_t smem = ((size_t)p->Nm * c + 32 + (size_t)ISLS_MAX_XROWS * c) * sizeof(double);
  if (S.P == 2 && S.c == 3 && S.ra == 4) k_sls_admm<2, 3, 4><<<(unsigned)B, threads, smem, s>>>(a, S, X);        // n = 4 (C4)
  else if (S.P == 2 && S.c == 2 && S.ra == 3) k_sls_admm<2, 2, 3><<<(unsigned)B, threads, smem, s>>>(a, S, X);   // n = 2
  else if (S.P == 2 && S.c == 4 && S.ra == 5) k_sls_admm<2, 4, 5><<<(unsigned)B, threads, smem, s>>>(a, S, X);   // n = 6
  else k_sls_admm<0, 0, 0><<<(unsigned)B, threads, smem, s>>>(a, S, X);
  CK(cudaGetLastError());
  return ISLS_OK;
}

// k_new = k + (I - K Su) Linv0 Su'Q (xd_new - xd_old)   (sls.py:244-248), one CTA per problem
__global__ void k_sls_replan(int Nm, int Nn, const double *Linv0, const double *DTQ, const double *Su, const double *K,
                             const double *k, const double *xdn, const double *xdo, double *kn) {
  extern __shared__ double sv[];            // [Nn] delta / Su t2, [Nm] t1, [Nm] t2
  double *dl = sv, *t1 = sv + Nn, *t2 = t1 + Nm;
  const long long b = blockIdx.x;
  for (int r = threadIdx.x; r < Nn; r += blockDim.x) dl[r] = xdn[(size_t)b * Nn + r] - xdo[(size_t)b * Nn + r];
  __syncthreads();
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < Nn; q++) acc = fma(DTQ[(size_t)r * Nn + q], dl[q], acc);
    t1[r] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < Nm; q++) acc = fma(Linv0[(size_t)r * Nm + q], t1[q], acc);
    t2[r] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Nn; r += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < Nm; q++) acc = fma(Su[(size_t)r * Nm + q], t2[q], acc);
    dl[r] = acc;
  }
  __syncthreads();
  const double *Kb = K + (size_t)b * Nm * Nn;
  for (int r = threadIdx.x; r < Nm; r += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < Nn; q++) acc = fma(Kb[(size_t)r * Nn + q], dl[q], acc);
    kn[(size_t)b * Nm + r] = k[(size_t)b * Nm + r] + (t2[r] - acc);
  }
}

extern "C" int isls_sls_replan_f64(const isls_sls_plan *p, int64_t B, const double *K_dev, const double *k_dev,
                                   const double *xd_new_dev, const double *xd_old_dev, double *k_new_dev, void *stream) {
  if (!p || B <= 0 || !K_dev || !k_dev || !xd_new_dev || !xd_old_dev || !k_new_dev)
    return isls_fail(ISLS_E_INVALID, "NULL argument or B <= 0");
  const size_t smem = ((size_t)p->Nn + 2 * p->Nm) * sizeof(double);
  k_sls_replan<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(p->Nm, p->Nn, p->Linv0, p->DTQ, p->Su, K_dev, k_dev,
                                                                xd_new_dev, xd_old_dev, k_new_dev);
  CK(cudaGetLastError());
  return ISLS_OK;
}

extern "C" size_t isls_controller_tv_workspace_bytes(int32_t n, int32_t m, int32_t N, int64_t B) {
  const size_t Nn = (size_t)N * n, Nm = (size_t)N * m;
  return (size_t)B * (2 * Nn * Nn + Nn * Nm) * sizeof(double);
}

// iSLS.controller for a general causal PHI_U on the time-varying operators (isls/sls.py:235-242 applied to the C, D of
// isls/isls_base.py:138-158): builds C, D of every problem from its A_t, B_t and runs the block back-substitution.
extern "C" int isls_controller_tv_f64(int32_t n, int32_t m, int32_t N, int64_t B, const double *A_dev,
                                      const double *B_dev, int32_t shared_AB, const double *PHI_U_dev,
                                      const double *du_dev, void *workspace_dev, size_t workspace_bytes,
                                      double *K_dev, double *k_dev, void *stream) {
  if (B <= 0 || N < 1 || n < 1 || m < 1 || !A_dev || !B_dev || !PHI_U_dev || !du_dev || !workspace_dev || !K_dev || !k_dev)
    return isls_fail(ISLS_E_INVALID, "NULL argument or bad size");
  if (n > TV_MAXN) return isls_fail(ISLS_E_UNSUPPORTED, "x_dim > 16");
  if (workspace_bytes < isls_controller_tv_workspace_bytes(n, m, N, B))
    return isls_fail(ISLS_E_WORKSPACE, "controller workspace too small (isls_controller_tv_workspace_bytes)");
  const size_t Nn = (size_t)N * n, Nm = (size_t)N * m;
  if (Nn * sizeof(double) > 48 * 1024) return isls_fail(ISLS_E_UNSUPPORTED, "N * x_dim > 6144");
  cudaStream_t s = (cudaStream_t)stream;
  double *C = (double *)workspace_dev, *D = C + (size_t)B * Nn * Nn, *PX = D + (size_t)B * Nn * Nm;
  const size_t as = shared_AB ? 0 : (size_t)N * n * n, bs = shared_AB ? 0 : (size_t)N * n * m;
  k_tv_build_CD<<<(unsigned)B, 256, 0, s>>>(n, m, N, A_dev, B_dev, as, bs, C, D);
  k_sls_controller<<<(unsigned)B, 256, Nn * sizeof(double), s>>>(n, m, N, (int)Nn, C, D, nullptr, PHI_U_dev, du_dev, PX, K_dev,
                                                                 k_dev, Nn * Nn, Nn * Nm);
  CK(cudaGetLastError());
  return ISLS_OK;
}

extern "C" int isls_sls_controller_f64(const isls_sls_plan *p, int64_t B, int32_t n_first_cols,
                                       const double *phi_cols_dev, const double *du_dev, void *workspace_dev,
                                       size_t workspace_bytes, double *K_dev, double *k_dev, void *stream) {
  if (!p || B <= 0 || !du_dev || !K_dev || !k_dev || !workspace_dev || n_first_cols < 0 || n_first_cols > p->Nn)
    return isls_fail(ISLS_E_INVALID, "NULL argument or bad size");
  if (n_first_cols > 0 && !phi_cols_dev) return isls_fail(ISLS_E_INVALID, "phi_cols_dev is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  static const bool dense = getenv("ISLS_SLS_CTRL_DENSE") != nullptr;       // test switch: per-problem dense form
  const size_t smem = ((size_t)p->Nn + (size_t)n_first_cols * p->Nm) * sizeof(double);
  if (dense || smem > 48 * 1024) {
    const size_t need = (size_t)B * p->Nn * p->Nn * sizeof(double);
    if (workspace_bytes < need) return isls_fail(ISLS_E_WORKSPACE, "controller workspace too small (B*(N n)^2 doubles)");
    k_sls_controller<<<(unsigned)B, 256, p->Nn * sizeof(double), s>>>(p->n, p->m, p->N, n_first_cols, p->Sw, p->Su, p->PHI,
                                                                     phi_cols_dev, du_dev, (double *)workspace_dev, K_dev,
                                                                     k_dev);
  } else {
    isls_sls_plan *pm = const_cast<isls_sls_plan *>(p);     // cache of the shared columns (one host thread per plan)
    if (pm->ctrl_cfirst != n_first_cols) {
      k_sls_ctrl_shared<<<1, 1024, 0, s>>>(p->n, p->m, p->N, n_first_cols, p->Sw, p->Su, p->PHI, pm->PXsh, pm->Ksh);
      pm->ctrl_cfirst = n_first_cols;
    }
    k_sls_ctrl_batch<<<(unsigned)B, 256, smem, s>>>(p->n, p->m, p->N, n_first_cols, p->Sw, p->Su, pm->Ksh, phi_cols_dev,
                                                    du_dev, K_dev, k_dev);
  }
  CK(cudaGetLastError());
  return ISLS_OK;
}
