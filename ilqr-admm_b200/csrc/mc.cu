// Monte-Carlo closed-loop evaluation of a controller over a batch of sampled initial states:
//   SLSBase / iSLSBase.get_trajectory_batch  (open loop us)              isls/sls_base.py:62-74,  isls/isls_base.py:45-58
//   SLSBase / iSLSBase.get_trajectory_dp     (u_t = K_t x_t + k_t)       isls/sls_base.py:76-89,  isls/isls_base.py:60-71
//   SLSBase / iSLSBase.get_trajectory_sls    (history feedback
//        u_t = sum_{s<=t} K[t,s] (x_s - x^_s) + k_t + u^_t)              isls/sls_base.py:91-105, isls/isls_base.py:28-43
// One sample per thread, the controller is shared by the batch (every lane reads the same gain -> broadcast loads).
// Process noise w ~ N(0, noise_scale) per state and step comes from a counter-based generator (Philox-4x32-10 +
// Box-Muller), so a run is reproducible from (seed, sample, step) - the reference uses the unseeded global numpy
// generator, i.e. only its statistics are defined.
#include <math.h>
#include <vector>
#include <stdint.h>

#include "../../include/isls_b200.h"
#include "common.cuh"
#include "models.cuh"
#include "soc.cuh"

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                           uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two standard normals from one Philox block (53-bit uniforms, Box-Muller)
__device__ __forceinline__ void normal2(unsigned long long seed, unsigned long long sample, uint32_t step, uint32_t pair,
                                        double &g0, double &g1) {
  uint32_t r[4];
  philox4x32((uint32_t)sample, (uint32_t)(sample >> 32), step, pair, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u0 = ((((unsigned long long)r[0] << 21) ^ (unsigned long long)(r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double u1 = ((((unsigned long long)r[2] << 21) ^ (unsigned long long)(r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u0));
  double s, c;
  sincos(6.283185307179586476925286766559 * u1, &s, &c);
  g0 = rad * c;
  g1 = rad * s;
}

#define MC_MAX_HIST 4608      // doubles of per-thread state history (N * n), local memory

template <class M>
__global__ void k_mc_rollout(int mode, int N, long long B, double dt, const double *x0, const double *K, const double *k,
                             const double *x_nom, const double *u_nom, double noise, unsigned long long seed,
                             double *x_out, double *u_out) {
  constexpr int n = M::n, m = M::m;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[n], u[m], xn[n];
  double hist[MC_MAX_HIST];             // (x_s - x^_s) history for the SLS feedback (mode 2 only)
#pragma unroll
  for (int i = 0; i < n; i++) x[i] = x0[b * n + i];
  for (int t = 0; t < N; t++) {
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < m; j++) u[j] = k[t * m + j];
    } else if (mode == 1) {
#pragma unroll
      for (int j = 0; j < m; j++) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) acc = fma(K[(size_t)(t * m + j) * n + i], x[i], acc);
        u[j] = acc + k[t * m + j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < n; i++) hist[t * n + i] = x[i] - (x_nom ? x_nom[t * n + i] : 0.0);
#pragma unroll
      for (int j = 0; j < m; j++) {
        const double *row = K + (size_t)(t * m + j) * N * n;
        double acc = 0.0;
        for (int q = 0; q < (t + 1) * n; q++) acc = fma(row[q], hist[q], acc);   // K is causal: columns > t are zero
        u[j] = (acc + k[t * m + j]) + (u_nom ? u_nom[t * m + j] : 0.0);
      }
    }
#pragma unroll
    for (int i = 0; i < n; i++) x_out[((size_t)b * N + t) * n + i] = x[i];
#pragma unroll
    for (int j = 0; j < m; j++) u_out[((size_t)b * N + t) * m + j] = u[j];
    M::step(x, u, xn, dt);
    if (noise != 0.0) {
#pragma unroll
      for (int i = 0; i < n; i += 2) {
        double g0, g1;
        normal2(seed, (unsigned long long)b, (uint32_t)t, (uint32_t)(i >> 1), g0, g1);
        xn[i] += noise * g0;
        if (i + 1 < n) xn[i + 1] += noise * g1;
      }
    }
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = xn[i];
  }
}

extern "C" int isls_mc_rollout_f64(int32_t model_id, int32_t n, int32_t m, int32_t N, double dt, int32_t mode,
                                   int64_t B, const double *x0_dev, const double *K_dev, const double *k_dev,
                                   const double *x_nom_dev, const double *u_nom_dev, double noise_scale, uint64_t seed,
                                   double *x_out_dev, double *u_out_dev, void *stream) {
  if (B <= 0 || N < 1 || !x0_dev || !k_dev || !x_out_dev || !u_out_dev) return isls_fail(ISLS_E_INVALID, "NULL argument or bad size");
  if (mode < 0 || mode > 2) return isls_fail(ISLS_E_INVALID, "mode must be 0 (open loop), 1 (dp) or 2 (sls)");
  if (mode >= 1 && !K_dev) return isls_fail(ISLS_E_INVALID, "K_dev is NULL");
  if (mode == 2 && (long long)N * n > MC_MAX_HIST) return isls_fail(ISLS_E_UNSUPPORTED, "N * x_dim too large for the SLS history");
  if (isls_model_supported(model_id, n, m)) return isls_fail(ISLS_E_UNSUPPORTED, "unsupported (model, n, m)");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((B + 127) / 128);
#define GO(MODEL) k_mc_rollout<MODEL><<<grid, 128, 0, s>>>(mode, N, B, dt, x0_dev, K_dev, k_dev, x_nom_dev, u_nom_dev, \
                                                          noise_scale, seed, x_out_dev, u_out_dev)
  if (model_id == ISLS_MODEL_CAR) GO(CarModel);
  else if (model_id == ISLS_MODEL_ARM3) GO(Arm3Model);
  else if (model_id == ISLS_MODEL_TASSA_CAR) GO(TassaCarModel);
  else if (m == 1) GO(DoubleIntModel<1>);
  else if (m == 2) GO(DoubleIntModel<2>);
  else GO(DoubleIntModel<3>);
#undef GO
  CK(cudaGetLastError());
  return ISLS_OK;
}

// ------------------------------------------------------------------------------------ get_AB of the notebooks
// A_t = df/dx, B_t = df/du of the registered model at (x_t, u_t) for every row of x [rows, n], u [rows, m]: the device
// form of the `get_AB(x_nom, u_nom)` callables the reference's callers pass (e.g. 3DoF robot notebooks cell 9,
// iLQR_ADMM.ipynb cell 5), in the natural layouts A [rows, n, n], B [rows, n, m] that iSLSBase.AB takes
// (isls/isls_base.py:133-158).  The solver kernels never need it (they linearise in registers); it feeds the
// stage-level API (iSLS.AB, iSLS.controller).
template <class M>
__global__ void k_linearize(long long rows, double dt, const double *x, const double *u, double *A, double *Bm) {
  constexpr int n = M::n, m = M::m;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double xv[n], uv[m], J[M::NJA], Am[n][n], Bv[n][m];
#pragma unroll
  for (int i = 0; i < n; i++) xv[i] = x[r * n + i];
#pragma unroll
  for (int j = 0; j < m; j++) uv[j] = u[r * m + j];
#pragma unroll
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int j = 0; j < n; j++) Am[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < m; j++) Bv[i][j] = 0.0;
  }
  M::jac(xv, uv, J, dt);
  M::expand(J, Am, Bv, dt);
#pragma unroll
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int j = 0; j < n; j++)      // entries the model declares structurally zero / one are not written by expand
      A[(r * n + i) * n + j] = M::am(i, j) == MZ ? 0.0 : M::am(i, j) == MO ? 1.0 : Am[i][j];
#pragma unroll
    for (int j = 0; j < m; j++) Bm[(r * n + i) * m + j] = M::bm(i, j) == MZ ? 0.0 : Bv[i][j];
  }
}

extern "C" int isls_linearize_f64(int32_t model_id, int32_t n, int32_t m, double dt, int64_t rows, const double *x_dev,
                                  const double *u_dev, double *A_dev, double *B_dev, void *stream) {
  if (rows <= 0 || !x_dev || !u_dev || !A_dev || !B_dev) return isls_fail(ISLS_E_INVALID, "NULL argument or rows <= 0");
  if (model_id == ISLS_MODEL_LTI) return isls_fail(ISLS_E_UNSUPPORTED, "an LTI model's A, B are its own constants");
  if (isls_model_supported(model_id, n, m)) return isls_fail(ISLS_E_UNSUPPORTED, "unsupported (model, n, m)");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows + 127) / 128);
#define GO(MODEL) k_linearize<MODEL><<<grid, 128, 0, s>>>(rows, dt, x_dev, u_dev, A_dev, B_dev)
  if (model_id == ISLS_MODEL_CAR) GO(CarModel);
  else if (model_id == ISLS_MODEL_ARM3) GO(Arm3Model);
  else if (model_id == ISLS_MODEL_TASSA_CAR) GO(TassaCarModel);
  else if (m == 1) GO(DoubleIntModel<1>);
  else if (m == 2) GO(DoubleIntModel<2>);
  else GO(DoubleIntModel<3>);
#undef GO
  CK(cudaGetLastError());
  return ISLS_OK;
}

// --------------------------------------------------------------------------- batched row projections
// Device counterparts of the `_batch` projections of isls/projections.py, one row per thread (dim <= 16):
//   kind 0 bound      np.clip(x, lo[dim], hi[dim])                                          projections.py:7-11
//   kind 1 linear     l <= a'x <= u                (a = p0[dim])                             projections.py:30-43
//   kind 2 quadratic  l <= 0.5 |x - c|^2 <= u      (c = p0[dim] or NULL)                     projections.py:86-104
//   kind 3 soc_unit   |z| <= t with x = [z, t], numpy batch semantics            projections.py:140-162
//   kind 4 square     l <= |x - c|_inf <= u        (c = p0[dim] or NULL)                     projections.py:252-272
//   kind 5 unit_ball  |x| <= 1                                                               projections.py:232-240
#define PROJ_MAXD 16
__global__ void k_project_rows(int kind, long long rows, int dim, const double *x, const double *p0, const double *p1,
                               double l, double u, double *out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double v[PROJ_MAXD], z[PROJ_MAXD];
  for (int i = 0; i < dim; i++) v[i] = x[r * dim + i];
  if (kind == 0) {
    for (int i = 0; i < dim; i++) z[i] = fmin(fmax(v[i], p0[i]), p1[i]);
  } else if (kind == 1) {
    double ax = 0.0, aa = 0.0;
    for (int i = 0; i < dim; i++) { ax += v[i] * p0[i]; aa += p0[i] * p0[i]; }
    aa += 1e-30;
    double mu = 0.0;
    if (ax > u) mu = ax - u;
    if (ax < l) mu = ax - l;
    for (int i = 0; i < dim; i++) z[i] = (mu != 0.0) ? v[i] - mu * (p0[i] / aa) : v[i];
  } else if (kind == 2) {
    double ss = 0.0;
    for (int i = 0; i < dim; i++) { if (p0) v[i] -= p0[i]; ss += v[i] * v[i]; }
    const double val = 0.5 * ss, nrm = sqrt(ss);
    for (int i = 0; i < dim; i++) z[i] = v[i];
    if (val > u) for (int i = 0; i < dim; i++) z[i] = v[i] * sqrt(2.0 * u) / nrm;
    if (l > val) for (int i = 0; i < dim; i++) z[i] = v[i] * sqrt(2.0 * l) / nrm;
    if (p0) for (int i = 0; i < dim; i++) z[i] += p0[i];
  } else if (kind == 3) {
    const int d = dim - 1;
    double ss = 0.0;
    for (int i = 0; i < d; i++) ss += v[i] * v[i];
    const double zn = sqrt(ss), t = v[d];
    const bool c1 = (zn <= -t) || (t < 0.0), c2 = (zn > t) || (zn > -t), c3 = zn <= t;
    for (int i = 0; i < dim; i++) z[i] = v[i];
    if (c2) { const double tmp = (zn + t) / 2.0; for (int i = 0; i < d; i++) z[i] = tmp * v[i] / (zn + 1e-30); z[d] = tmp; }
    if (c1) for (int i = 0; i < dim; i++) z[i] = 0.0;
    if (c3) for (int i = 0; i < dim; i++) z[i] = v[i];
  } else if (kind == 4) {
    int j = 0;
    double mx = -1.0;
    for (int i = 0; i < dim; i++) { if (p0) v[i] -= p0[i]; const double a = fabs(v[i]); if (a > mx) { mx = a; j = i; } }
    for (int i = 0; i < dim; i++) z[i] = v[i];
    if (mx < l) z[j] = l * ((v[j] > 0.0) - (v[j] < 0.0));
    for (int i = 0; i < dim; i++) { z[i] = fmax(fmin(z[i], u), -u); if (p0) z[i] += p0[i]; }
  } else {
    double ss = 0.0;
    for (int i = 0; i < dim; i++) ss += v[i] * v[i];
    const double nrm = sqrt(ss);
    for (int i = 0; i < dim; i++) z[i] = (nrm <= 1.0) ? v[i] : v[i] / nrm;
  }
  for (int i = 0; i < dim; i++) out[r * dim + i] = z[i];
}

extern "C" int isls_project_rows_f64(int32_t kind, int64_t rows, int32_t dim, const double *x_dev, const double *p0_dev,
                                     const double *p1_dev, double l, double u, double *out_dev, void *stream) {
  if (rows <= 0 || dim < 1 || dim > PROJ_MAXD || !x_dev || !out_dev) return isls_fail(ISLS_E_INVALID, "bad size or NULL argument");
  if (kind < 0 || kind > 5) return isls_fail(ISLS_E_UNSUPPORTED, "unknown projection kind");
  if ((kind == 0 && (!p0_dev || !p1_dev)) || (kind == 1 && !p0_dev)) return isls_fail(ISLS_E_INVALID, "missing parameter array");
  if (kind == 3 && dim < 2) return isls_fail(ISLS_E_INVALID, "soc_unit needs dim >= 2");
  k_project_rows<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(kind, rows, dim, x_dev, p0_dev, p1_dev, l, u,
                                                                                   out_dev);
  CK(cudaGetLastError());
  return ISLS_OK;
}

// ------------------------------------------------------------------ batched row projections, parameterised (8f #2)
// project_multilinear (isls/projections.py:46-62), project_soc (163-232: general cone A z + b in SOC by an inner ADMM whose
// stop rule is a maximum over ALL rows - one CTA, one row per thread), project_block_lower_triangular (277-286).
#define PROJ_MAXK 8
struct ProjEx {
  int k, dim, max_iter;
  double A[PROJ_MAXK][PROJ_MAXD];
  double b[PROJ_MAXK], l[PROJ_MAXK], u[PROJ_MAXK];
  double inv[PROJ_MAXD][PROJ_MAXD];      // multilinear: (A A')^-1 [k, k]; soc: (I + rho A'A)^-1 [dim, dim]
  double rho, tol;
};

__global__ void k_project_multilinear(ProjEx P, long long rows, const double *x, double *out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double v[PROJ_MAXD], d[PROJ_MAXK], mu[PROJ_MAXK];
  for (int i = 0; i < P.dim; i++) v[i] = x[r * P.dim + i];
  for (int a = 0; a < P.k; a++) {
    double ax = 0.0;
    for (int i = 0; i < P.dim; i++) ax += P.A[a][i] * v[i];
    double t = ax;
    if (ax > P.u[a]) t = P.u[a];
    if (ax < P.l[a]) t = P.l[a];                       // the later mask wins, like the numpy code
    d[a] = ax - t;
  }
  for (int a = 0; a < P.k; a++) {
    double m = 0.0;
    for (int c = 0; c < P.k; c++) m += P.inv[a][c] * d[c];
    mu[a] = m;
  }
  for (int i = 0; i < P.dim; i++) {
    double c = 0.0;
    for (int a = 0; a < P.k; a++) c += P.A[a][i] * mu[a];
    out[r * P.dim + i] = v[i] - c;
  }
}

__global__ void k_project_soc(ProjEx P, int rows, const double *x, double *out, int *iters) {
  __shared__ double red[32];
  const int r = threadIdx.x;
  const bool act = r < rows;
  const int n = P.dim, k = P.k;
  double z0[PROJ_MAXD], z[PROJ_MAXD], xs[PROJ_MAXK], lm[PROJ_MAXK], y[PROJ_MAXK], azb[PROJ_MAXK];
  for (int i = 0; i < n; i++) { z0[i] = act ? x[(size_t)r * n + i] : 0.0; z[i] = z0[i]; }
  for (int a = 0; a < k; a++) lm[a] = 0.0;
  double pm = 1e5, dm = 1e5;
  int it = 0;
  for (int j = 0; j < P.max_iter; j++) {
    it = j + 1;
    for (int a = 0; a < k; a++) {
      double s = P.b[a];
      for (int i = 0; i < n; i++) s += P.A[a][i] * z[i];
      y[a] = s + lm[a];
    }
    soc_unit_row(k - 1, y, xs);                           // x = project_soc_unit(A z + b + lambda)
    double zp[PROJ_MAXD], rs[PROJ_MAXD];
    for (int i = 0; i < n; i++) {
      zp[i] = z[i];
      double s = 0.0;
      for (int a = 0; a < k; a++) s += P.A[a][i] * ((-P.b[a] + xs[a]) - lm[a]);
      rs[i] = z0[i] + P.rho * s;
    }
    for (int i = 0; i < n; i++) {
      double s = 0.0;
      for (int c = 0; c < n; c++) s += P.inv[i][c] * rs[c];
      z[i] = s;
    }
    double ps = 0.0, ds = 0.0;
    for (int a = 0; a < k; a++) {
      double s = P.b[a];
      for (int i = 0; i < n; i++) s += P.A[a][i] * z[i];
      azb[a] = s;
      const double pr = s - xs[a];
      ps += pr * pr;
      lm[a] += pr;
    }
    for (int i = 0; i < n; i++) { const double dv = P.rho * (z[i] - zp[i]); ds += dv * dv; }
    const double pprev = pm, dprev = dm;
    pm = block_max(act ? sqrt(ps) : 0.0, red);
    dm = block_max(act ? sqrt(ds) : 0.0, red);
    if (pm < P.tol && dm < P.tol) break;
    if (j < P.max_iter - 1) {
      const double pc = fabs(pprev - pm) / (pprev + 1e-30), dc = fabs(dprev - dm) / (dprev + 1e-30);
      if (pc < 1e-5 && dc < 1e-5) break;
    }
  }
  if (act)
    for (int i = 0; i < n; i++) out[(size_t)r * n + i] = z[i];
  if (iters && r == 0) iters[0] = it;
}

__global__ void k_project_blt(double *z, int x_dim, int u_dim, int N) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= N * x_dim) return;
  const int i = q / x_dim, c = q % x_dim;
  z[(size_t)(i * u_dim) * ((size_t)N * x_dim) + (size_t)i * x_dim + c] = 0.0;
}

static bool gauss_jordan_inv(int n, const double *Min, double (*out)[PROJ_MAXD]) {
  double Mx[PROJ_MAXD][2 * PROJ_MAXD] = {};
  for (int a = 0; a < n; a++)
    for (int c = 0; c < n; c++) { Mx[a][c] = Min[a * n + c]; Mx[a][n + c] = a == c ? 1.0 : 0.0; }
  for (int col = 0; col < n; col++) {
    int piv = col;
    for (int r2 = col + 1; r2 < n; r2++) if (fabs(Mx[r2][col]) > fabs(Mx[piv][col])) piv = r2;
    if (Mx[piv][col] == 0.0) return false;
    for (int q = 0; q < 2 * n; q++) std::swap(Mx[col][q], Mx[piv][q]);
    const double d = Mx[col][col];
    for (int q = 0; q < 2 * n; q++) Mx[col][q] /= d;
    for (int r2 = 0; r2 < n; r2++)
      if (r2 != col) {
        const double f = Mx[r2][col];
        for (int q = 0; q < 2 * n; q++) Mx[r2][q] -= f * Mx[col][q];
      }
  }
  for (int a = 0; a < n; a++)
    for (int c = 0; c < n; c++) out[a][c] = Mx[a][n + c];
  return true;
}

extern "C" int isls_project_rows_ex_f64(const isls_proj_params *pp, int64_t rows, int32_t dim, const double *x_dev,
                                        double *out_dev, int32_t *iters_dev, void *stream) {
  if (!pp) return isls_fail(ISLS_E_INVALID, "params is NULL");
  if (pp->struct_size != (uint32_t)sizeof(isls_proj_params))
    return isls_fail(ISLS_E_INVALID, "isls_proj_params.struct_size mismatch (binding built against another isls_b200.h?)");
  cudaStream_t s = (cudaStream_t)stream;
  if (pp->kind == ISLS_PROJ_BLOCK_LOWER_TRIANGULAR) {
    if (!out_dev || pp->x_dim < 1 || pp->u_dim < 1 || pp->N < 1) return isls_fail(ISLS_E_INVALID, "bad size or NULL argument");
    const int tot = pp->N * pp->x_dim;
    k_project_blt<<<(tot + 127) / 128, 128, 0, s>>>(out_dev, pp->x_dim, pp->u_dim, pp->N);
    CK(cudaGetLastError());
    return ISLS_OK;
  }
  if (rows <= 0 || dim < 1 || dim > PROJ_MAXD || !x_dev || !out_dev) return isls_fail(ISLS_E_INVALID, "bad size or NULL argument");
  if (pp->k < 1 || pp->k > PROJ_MAXK || !pp->A) return isls_fail(ISLS_E_INVALID, "A missing or k out of range (1..8)");
  ProjEx P;
  memset(&P, 0, sizeof(P));
  P.k = pp->k; P.dim = dim; P.max_iter = pp->max_iter; P.rho = pp->rho; P.tol = pp->tol;
  for (int a = 0; a < P.k; a++) {
    for (int i = 0; i < dim; i++) P.A[a][i] = pp->A[a * dim + i];
    P.b[a] = pp->b ? pp->b[a] : 0.0;
    P.l[a] = pp->l ? pp->l[a] : -INFINITY;
    P.u[a] = pp->u ? pp->u[a] : INFINITY;
  }
  if (pp->kind == ISLS_PROJ_MULTILINEAR) {
    double AAT[PROJ_MAXK * PROJ_MAXK];
    for (int a = 0; a < P.k; a++)
      for (int c = 0; c < P.k; c++) {
        double v = 0.0;
        for (int i = 0; i < dim; i++) v += P.A[a][i] * P.A[c][i];
        AAT[a * P.k + c] = v;
      }
    if (!gauss_jordan_inv(P.k, AAT, P.inv)) return isls_fail(ISLS_E_INVALID, "A A' is singular");
    k_project_multilinear<<<(unsigned)((rows + 127) / 128), 128, 0, s>>>(P, rows, x_dev, out_dev);
  } else if (pp->kind == ISLS_PROJ_SOC) {
    if (rows > 1024) return isls_fail(ISLS_E_UNSUPPORTED, "project_soc: the stop rule couples all rows; rows <= 1024");
    if (P.k < 2 || P.max_iter < 1) return isls_fail(ISLS_E_INVALID, "project_soc needs k >= 2 and max_iter >= 1");
    double Lm[PROJ_MAXD * PROJ_MAXD];
    for (int i = 0; i < dim; i++)
      for (int c = 0; c < dim; c++) {
        double v = i == c ? 1.0 : 0.0;
        for (int a = 0; a < P.k; a++) v += P.rho * P.A[a][i] * P.A[a][c];
        Lm[i * dim + c] = v;
      }
    if (!gauss_jordan_inv(dim, Lm, P.inv)) return isls_fail(ISLS_E_INVALID, "I + rho A'A is singular");
    k_project_soc<<<1, (unsigned)(((rows + 31) / 32) * 32), 0, s>>>(P, (int)rows, x_dev, out_dev, iters_dev);
  } else {
    return isls_fail(ISLS_E_UNSUPPORTED, "unknown projection kind");
  }
  CK(cudaGetLastError());
  return ISLS_OK;
}

// ---- generic project_set_convex (isls/projections.py:289-374) over up to 4 sets {A_i x + b_i in C_i}, each C_i one of the
// primitive row projections (bound / quadratic shell / SOC / infinity-norm shell / unit ball, batch semantics), all rows
// projected together by one CTA (the stop rule is a maximum over sets and rows).
#define PSC_MAXS 4
struct ProjSet {
  int K, dim, max_iter;
  int ra[PSC_MAXS], kind[PSC_MAXS];
  double A[PSC_MAXS][PROJ_MAXK][PROJ_MAXD], b[PSC_MAXS][PROJ_MAXK];
  double p0[PSC_MAXS][PROJ_MAXK], p1[PSC_MAXS][PROJ_MAXK], l[PSC_MAXS], u[PSC_MAXS];
  int has_p0[PSC_MAXS];
  double inv[PROJ_MAXD][PROJ_MAXD];
  double rho, thr;
};
__device__ static void proj_primitive(int kind, int d, const double *v_in, const double *p0, const double *p1, int has_p0,
                                      double l, double u, double *z) {
  double v[PROJ_MAXK];
  for (int i = 0; i < d; i++) v[i] = v_in[i];
  if (kind == 0) {
    for (int i = 0; i < d; i++) z[i] = fmin(fmax(v[i], p0[i]), p1[i]);
  } else if (kind == 2) {
    double ss = 0.0;
    for (int i = 0; i < d; i++) { if (has_p0) v[i] -= p0[i]; ss += v[i] * v[i]; }
    const double val = 0.5 * ss, nrm = sqrt(ss);
    for (int i = 0; i < d; i++) z[i] = v[i];
    if (val > u) for (int i = 0; i < d; i++) z[i] = v[i] * sqrt(2.0 * u) / nrm;
    if (l > val) for (int i = 0; i < d; i++) z[i] = v[i] * sqrt(2.0 * l) / nrm;
    if (has_p0) for (int i = 0; i < d; i++) z[i] += p0[i];
  } else if (kind == 3) {
    soc_unit_row(d - 1, v, z);
  } else if (kind == 4) {
    int j = 0;
    double mx = -1.0;
    for (int i = 0; i < d; i++) { if (has_p0) v[i] -= p0[i]; const double a = fabs(v[i]); if (a > mx) { mx = a; j = i; } }
    for (int i = 0; i < d; i++) z[i] = v[i];
    if (mx < l) z[j] = l * ((v[j] > 0.0) - (v[j] < 0.0));
    for (int i = 0; i < d; i++) { z[i] = fmax(fmin(z[i], u), -u); if (has_p0) z[i] += p0[i]; }
  } else {
    double ss = 0.0;
    for (int i = 0; i < d; i++) ss += v[i] * v[i];
    const double nrm = sqrt(ss);
    for (int i = 0; i < d; i++) z[i] = (nrm <= 1.0) ? v[i] : v[i] / nrm;
  }
}

__global__ void k_project_set_convex(ProjSet P, int rows, const double *x, double *out, int *iters) {
  __shared__ double red[32];
  const int r = threadIdx.x, n = P.dim;
  const bool act = r < rows;
  double x0[PROJ_MAXD], xv[PROJ_MAXD], zi[PSC_MAXS][PROJ_MAXK], li[PSC_MAXS][PROJ_MAXK];
  for (int i = 0; i < n; i++) { x0[i] = act ? x[(size_t)r * n + i] : 0.0; xv[i] = x0[i]; }
  for (int s = 0; s < P.K; s++)
    for (int e = 0; e < P.ra[s]; e++) {
      double v = P.b[s][e];
      for (int i = 0; i < n; i++) v += P.A[s][e][i] * xv[i];
      zi[s][e] = v;
      li[s][e] = 0.0;
    }
  double pm = 1e5, dm = 1e5;
  int it = 0;
  for (int j = 0; j < P.max_iter; j++) {
    it = j + 1;
    double rs[PROJ_MAXD];
    for (int i = 0; i < n; i++) rs[i] = 0.0;
    for (int s = 0; s < P.K; s++)
      for (int e = 0; e < P.ra[s]; e++) {
        const double w = (-P.b[s][e] + zi[s][e]) - li[s][e];
        for (int i = 0; i < n; i++) rs[i] += P.A[s][e][i] * w;
      }
    for (int i = 0; i < n; i++) rs[i] = x0[i] + P.rho * rs[i];
    for (int i = 0; i < n; i++) {
      double v = 0.0;
      for (int c = 0; c < n; c++) v += P.inv[i][c] * rs[c];
      xv[i] = v;
    }
    double pmax = 0.0, dmax = 0.0;
    for (int s = 0; s < P.K; s++) {
      double axb[PROJ_MAXK], y[PROJ_MAXK], zn[PROJ_MAXK];
      for (int e = 0; e < P.ra[s]; e++) {
        double v = P.b[s][e];
        for (int i = 0; i < n; i++) v += P.A[s][e][i] * xv[i];
        axb[e] = v;
        y[e] = v + li[s][e];
      }
      proj_primitive(P.kind[s], P.ra[s], y, P.p0[s], P.p1[s], P.has_p0[s], P.l[s], P.u[s], zn);
      double ps = 0.0, dr[PROJ_MAXD];
      for (int i = 0; i < n; i++) dr[i] = 0.0;
      for (int e = 0; e < P.ra[s]; e++) {
        const double pr = axb[e] - zn[e], dz = zn[e] - zi[s][e];
        ps += pr * pr;
        for (int i = 0; i < n; i++) dr[i] += P.A[s][e][i] * dz;
        li[s][e] += pr;
        zi[s][e] = zn[e];
      }
      double ds = 0.0;
      for (int i = 0; i < n; i++) ds += (P.rho * dr[i]) * (P.rho * dr[i]);
      pmax = fmax(pmax, sqrt(ps));
      dmax = fmax(dmax, sqrt(ds));
    }
    const double pprev = pm, dprev = dm;
    pm = block_max(act ? pmax : 0.0, red);
    dm = block_max(act ? dmax : 0.0, red);
    if (pm < P.thr && dm < P.thr) break;
    if (j < P.max_iter - 1) {
      const double pc = fabs(pprev - pm) / (pprev + 1e-30), dc = fabs(dprev - dm) / (dprev + 1e-30);
      if (pc < 1e-5 && dc < 1e-5) break;
    }
  }
  if (act)
    for (int i = 0; i < n; i++) out[(size_t)r * n + i] = xv[i];
  if (iters && r == 0) iters[0] = it;
}

extern "C" int isls_project_set_convex_f64(const isls_proj_set_params *pp, int64_t rows, int32_t dim, const double *x_dev,
                                           double *out_dev, int32_t *iters_dev, void *stream) {
  if (!pp || rows <= 0 || rows > 1024 || dim < 1 || dim > PROJ_MAXD || !x_dev || !out_dev)
    return isls_fail(ISLS_E_INVALID, "bad size (rows <= 1024, dim <= 16) or NULL argument");
  if (pp->struct_size != (uint32_t)sizeof(isls_proj_set_params))
    return isls_fail(ISLS_E_INVALID, "isls_proj_set_params.struct_size mismatch (binding built against another isls_b200.h?)");
  if (pp->n_sets < 1 || pp->n_sets > PSC_MAXS || pp->max_iter < 1) return isls_fail(ISLS_E_INVALID, "n_sets in 1..4, max_iter >= 1");
  std::vector<ProjSet> holder(1);   // ~7 KB: kept off the caller's stack, no static state (the library is re-entrant)
  ProjSet &P = holder[0];
  memset(&P, 0, sizeof(P));
  P.K = pp->n_sets; P.dim = dim; P.max_iter = pp->max_iter; P.rho = pp->rho; P.thr = pp->threshold;
  for (int s = 0; s < P.K; s++) {
    const isls_proj_set_entry &e = pp->sets[s];
    if (e.rows < 1 || e.rows > PROJ_MAXK || !e.A) return isls_fail(ISLS_E_INVALID, "set: A missing or rows out of range (1..8)");
    if (!(e.kind == 0 || e.kind == 2 || e.kind == 3 || e.kind == 4 || e.kind == 5))
      return isls_fail(ISLS_E_UNSUPPORTED, "set kind must be bound (0), quadratic (2), soc_unit (3), square (4) or unit_ball (5)");
    if (e.kind == 0 && (!e.p0 || !e.p1)) return isls_fail(ISLS_E_INVALID, "bound set needs p0 = lo, p1 = hi");
    P.ra[s] = e.rows; P.kind[s] = e.kind; P.l[s] = e.l; P.u[s] = e.u; P.has_p0[s] = e.p0 != nullptr;
    for (int a = 0; a < e.rows; a++) {
      for (int i = 0; i < dim; i++) P.A[s][a][i] = e.A[a * dim + i];
      P.b[s][a] = e.b ? e.b[a] : 0.0;
      P.p0[s][a] = e.p0 ? e.p0[a] : 0.0;
      P.p1[s][a] = e.p1 ? e.p1[a] : 0.0;
    }
  }
  double Lm[PROJ_MAXD * PROJ_MAXD];
  for (int i = 0; i < dim; i++)
    for (int c = 0; c < dim; c++) {
      double v = i == c ? 1.0 : 0.0;
      for (int s = 0; s < P.K; s++)
        for (int a = 0; a < P.ra[s]; a++) v += P.rho * P.A[s][a][i] * P.A[s][a][c];
      Lm[i * dim + c] = v;
    }
  if (!gauss_jordan_inv(dim, Lm, P.inv)) return isls_fail(ISLS_E_INVALID, "I + rho sum A'A is singular");
  k_project_set_convex<<<1, (unsigned)(((rows + 31) / 32) * 32), 0, (cudaStream_t)stream>>>(P, (int)rows, x_dev, out_dev, iters_dev);
  CK(cudaGetLastError());
  return ISLS_OK;
}
