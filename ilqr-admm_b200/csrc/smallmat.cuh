// Small dense FP64 algebra for one trajectory per thread.  All loops are fully unrolled; the model's compile-time
// sparsity classes (models.cuh: MZ / MO / MV) remove structural zeros and turn structural ones into adds, so the
// generic recursion is specialised per model without hand-written variants.
#pragma once
#include "models.cuh"

// y = A^T v
template <class M>
__device__ __forceinline__ void mat_At_v(const double (&A)[M::n][M::n], const double (&v)[M::n],
                                         double (&y)[M::n]) {
#pragma unroll
  for (int j = 0; j < M::n; j++) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < M::n; k++) {
      if (M::am(k, j) == MO) acc += v[k];
      else if (M::am(k, j) == MV) acc = fma(A[k][j], v[k], acc);
    }
    y[j] = acc;
  }
}

// y = B^T v
template <class M>
__device__ __forceinline__ void mat_Bt_v(const double (&B)[M::n][M::m], const double (&v)[M::n],
                                         double (&y)[M::m]) {
#pragma unroll
  for (int j = 0; j < M::m; j++) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < M::n; k++) {
      if (M::bm(k, j) == MO) acc += v[k];
      else if (M::bm(k, j) == MV) acc = fma(B[k][j], v[k], acc);
    }
    y[j] = acc;
  }
}

// y = A x + B u
template <class M>
__device__ __forceinline__ void mat_Ax_Bu(const double (&A)[M::n][M::n], const double (&B)[M::n][M::m],
                                          const double (&x)[M::n], const double (&u)[M::m], double (&y)[M::n]) {
#pragma unroll
  for (int i = 0; i < M::n; i++) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < M::n; k++) {
      if (M::am(i, k) == MO) acc += x[k];
      else if (M::am(i, k) == MV) acc = fma(A[i][k], x[k], acc);
    }
#pragma unroll
    for (int k = 0; k < M::m; k++) {
      if (M::bm(i, k) == MO) acc += u[k];
      else if (M::bm(i, k) == MV) acc = fma(B[i][k], u[k], acc);
    }
    y[i] = acc;
  }
}

// VA = V A   (n x n)
template <class M>
__device__ __forceinline__ void mat_V_A(const double (&V)[M::n][M::n], const double (&A)[M::n][M::n],
                                        double (&VA)[M::n][M::n]) {
#pragma unroll
  for (int i = 0; i < M::n; i++)
#pragma unroll
    for (int j = 0; j < M::n; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < M::n; k++) {
        if (M::am(k, j) == MO) acc += V[i][k];
        else if (M::am(k, j) == MV) acc = fma(V[i][k], A[k][j], acc);
      }
      VA[i][j] = acc;
    }
}

// VB = V B   (n x m)
template <class M>
__device__ __forceinline__ void mat_V_B(const double (&V)[M::n][M::n], const double (&B)[M::n][M::m],
                                        double (&VB)[M::n][M::m]) {
#pragma unroll
  for (int i = 0; i < M::n; i++)
#pragma unroll
    for (int j = 0; j < M::m; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < M::n; k++) {
        if (M::bm(k, j) == MO) acc += V[i][k];
        else if (M::bm(k, j) == MV) acc = fma(V[i][k], B[k][j], acc);
      }
      VB[i][j] = acc;
    }
}

// C = A^T X  (n x P) for X (n x P)
template <class M, int P>
__device__ __forceinline__ void mat_At_X(const double (&A)[M::n][M::n], const double (&X)[M::n][P],
                                         double (&C)[M::n][P]) {
#pragma unroll
  for (int i = 0; i < M::n; i++)
#pragma unroll
    for (int j = 0; j < P; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < M::n; k++) {
        if (M::am(k, i) == MO) acc += X[k][j];
        else if (M::am(k, i) == MV) acc = fma(A[k][i], X[k][j], acc);
      }
      C[i][j] = acc;
    }
}

// C = B^T X  (m x P) for X (n x P)
template <class M, int P>
__device__ __forceinline__ void mat_Bt_X(const double (&B)[M::n][M::m], const double (&X)[M::n][P],
                                         double (&C)[M::m][P]) {
#pragma unroll
  for (int i = 0; i < M::m; i++)
#pragma unroll
    for (int j = 0; j < P; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < M::n; k++) {
        if (M::bm(k, i) == MO) acc += X[k][j];
        else if (M::bm(k, i) == MV) acc = fma(B[k][i], X[k][j], acc);
      }
      C[i][j] = acc;
    }
}

// Cholesky-based inverse of a small SPD matrix.  Returns false if a pivot is not strictly positive (the
// reference's dposv raises LinAlgError there, isls/isls.py:296); the inverse is then left as identity.
template <int m>
__device__ __forceinline__ bool spd_inverse(const double (&Q)[m][m], double (&Qi)[m][m]) {
  double Lc[m][m];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < m; j++) {
    double s = Q[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) s -= Lc[j][k] * Lc[j][k];
    if (!(s > 0.0)) { ok = false; s = 1.0; }
    const double d = sqrt(s);
    Lc[j][j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < m; i++) {
      double a = Q[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) a -= Lc[i][k] * Lc[j][k];
      Lc[i][j] = a * inv;
    }
  }
  // Li = L^-1 (lower)
  double Li[m][m];
#pragma unroll
  for (int j = 0; j < m; j++) {
#pragma unroll
    for (int i = 0; i < m; i++) {
      if (i < j) { Li[i][j] = 0.0; continue; }
      double a = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = j; k < i; k++) a -= Lc[i][k] * Li[k][j];
      Li[i][j] = a / Lc[i][i];
    }
  }
  // Qi = Li^T Li
#pragma unroll
  for (int i = 0; i < m; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      double a = 0.0;
#pragma unroll
      for (int k = i; k < m; k++) a = fma(Li[k][i], Li[k][j], a);
      Qi[i][j] = a;
      Qi[j][i] = a;
    }
  if (!ok) {
#pragma unroll
    for (int i = 0; i < m; i++)
#pragma unroll
      for (int j = 0; j < m; j++) Qi[i][j] = (i == j) ? 1.0 : 0.0;
  }
  return ok;
}
