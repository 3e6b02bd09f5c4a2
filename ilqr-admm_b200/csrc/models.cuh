// Device-side dynamics models registered behind the reference's forward_model / get_AB plugin slots.
//
// Every model provides
//   n, m          state / control dimension
//   NJ            number of trajectory-dependent Jacobian scalars cached per time step
//   am(i,j), bm(i,j)   compile-time sparsity class of A[i][j], B[i][j]: MZ (structural 0), MO (structural 1),
//                      MV (value held in the dense array).  The small-matrix helpers skip MZ terms and turn MO
//                      terms into plain adds, so one generic Riccati/ff/rollout code path is specialised per model
//                      by full unrolling.
//   step(x,u,xn,dt)    x_{t+1} = f(x_t,u_t)                 (reference: forward_model callable)
//   jac(x,u,J,dt)      the NJ scalars at (x_t,u_t)          (reference: get_AB callable)
//   expand(J,A,B,dt)   fill the MV entries of dense A, B
//
// Formulas (cited in oracle/models.py as well):
//   car   notebooks/Car/Iterative LQR with control constraints.ipynb cell 6
//   arm3  notebooks/3DoF robot/State and control bound constraints.ipynb cells 9-10, closed-form planar 3R
//         FK / Jacobian for urdfs/3dof_robot.urdf:73-102 (unit links)
//   double integrator   isls/utils.py:266-276, isls/sls_base.py:49-53
//   tassa_car   notebooks/Tutorial.ipynb cell 8
#pragma once

enum { MZ = 0, MO = 1, MV = 2 };

#define ISLS_TWO_PI 6.283185307179586476925286766559

static __device__ __noinline__ double mod_two_pi_slow(double a) {
  double r = fmod(a, ISLS_TWO_PI);
  if (r < 0.0) r += ISLS_TWO_PI;
  else if (r == 0.0) r = 0.0;                     // copysign(0, b) = +0
  return r;
}
// numpy's np.mod(a, 2*pi) for float64: r = fmod(a,b); if r != 0 and sign differs from b: r += b.
// Branch-free on [-2pi, 4pi) where it is exact by construction: [0,2pi) -> a; [2pi,4pi) -> a - 2pi (exact,
// Sterbenz) = fmod; [-2pi,0) -> fmod gives a, then a + 2pi (one rounding, as numpy).  Outside: `bad`.
__device__ __forceinline__ double mod_two_pi_core(double a, bool &bad) {
  double r = a;
  r = (a >= ISLS_TWO_PI) ? a - ISLS_TWO_PI : r;
  r = (a < 0.0) ? a + ISLS_TWO_PI : r;
  bad |= !(a >= -ISLS_TWO_PI && a < 2.0 * ISLS_TWO_PI);
  return r;
}
__device__ __forceinline__ double mod_two_pi(double a) {
  bool bad = false;
  const double r = mod_two_pi_core(a, bad);
  return bad ? mod_two_pi_slow(a) : r;
}

// sin and cos of a double with a 3-term Cody-Waite reduction by pi/2 (exact products through FMA) and the
// fdlibm kernel polynomials on [-pi/4, pi/4] (< 1 ulp).  The rollouts evaluate it once per candidate-step.  The
// coefficients live in constant memory so every DFMA takes its coefficient as a constant-bank operand (literal
// doubles cost two UMOV each per use - measured 25 % of the line-search instruction stream); |x| >= 2^17 takes the
// out-of-line library path.
static __constant__ double kSC[16] = {
    0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

static __device__ __noinline__ void sincos_slow(double x, double *sp, double *cp) { sincos(x, sp, cp); }

// branch-free core: `bad` is OR-ed with "argument outside the fast range" (results are then garbage and the caller
// must redo with sincos_slow); no branch, so several independent evaluations schedule into one basic block.
__device__ __forceinline__ void sincos_core(double x, double *sp, double *cp, bool &bad) {
  // k = rint(x * 2/pi) by the 1.5 * 2^52 magic-number trick: one FMA + one ADD on the FP64 pipe and the quadrant in
  // the low word of the biased sum (no FRND / F2I on the conversion pipe)
  const double big = fma(x, kSC[0], 6755399441055744.0);
  const double kd = big - 6755399441055744.0;
  const int k = __double2loint(big);
  bad |= !(fabs(x) < 200000.0);
  double r = fma(-kd, kSC[1], x);
  r = fma(-kd, kSC[2], r);
  r = fma(-kd, kSC[3], r);
  const double z = r * r;
  double ps = fma(z, kSC[4], kSC[5]);
  ps = fma(z, ps, kSC[6]);
  ps = fma(z, ps, kSC[7]);
  ps = fma(z, ps, kSC[8]);
  ps = fma(z, ps, kSC[9]);
  const double s = fma(z * r, ps, r);
  double pc = fma(z, kSC[10], kSC[11]);
  pc = fma(z, pc, kSC[12]);
  pc = fma(z, pc, kSC[13]);
  pc = fma(z, pc, kSC[14]);
  pc = fma(z, pc, kSC[15]);
  pc = fma(z, pc, -0.5);
  const double c = fma(z, pc, 1.0);          // 1 - z/2 + z^2 (C1 + ...), max error 1 ulp on [-pi/4, pi/4]
  const double sa = (k & 1) ? c : s, ca = (k & 1) ? s : c;
  *sp = (k & 2) ? -sa : sa;
  *cp = ((k + 1) & 2) ? -ca : ca;
}

// The slow path returns by value: handing the caller's s / c to a non-inlined function by address gave them a home in
// local memory (STL in the fast path, LDL after the reconvergence point - an L1-missing round trip per model
// evaluation in the cp.async-staged kernels, whose streams flush L1; profiles/r1_c3_small_batch_kernels.md).
static __device__ __noinline__ double2 sincos_slow2(double x) {
  double s, c;
  sincos(x, &s, &c);
  return make_double2(s, c);
}
__device__ __forceinline__ void sincos_pio2(double x, double *sp, double *cp) {
  bool bad = false;
  double s, c;
  sincos_core(x, &s, &c, bad);
  if (bad) {
    const double2 r = sincos_slow2(x);
    s = r.x;
    c = r.y;
  }
  *sp = s;
  *cp = c;
}

// ---- K independent evaluations written STAGE BY STAGE (loop over the K arguments inside every polynomial stage).
// ptxas keeps roughly the source order under register pressure: with the chain-by-chain form above the K Horner
// chains of a line-search thread were emitted one after the other (7 dependent DFMAs, 8 stall cycles each); in this
// form consecutive instructions belong to different chains, so a dependent DFMA is 2K issue slots away (>= the 8-cycle
// FP64 latency for K >= 4) and one warp alone can keep the FP64 pipe busy.  Same arithmetic as sincos_core, bit for bit.
// Sign flips are integer XORs on the high word (the chain form spent a DADD + 2 FSEL each).
__device__ __forceinline__ double xor_sign(double v, int w) {
  return __hiloint2double(__double2hiint(v) ^ (w & (int)0x80000000), __double2loint(v));
}
// |x| < 200000 on the high word only (no FP64-pipe compare); NaN / inf are out of range
__device__ __forceinline__ bool sincos_fast_range(double x) {
  return (__double2hiint(x) & 0x7fffffff) < 0x41086a00;          // 0x41086a00'00000000 = 200000.0
}
template <int K, bool CHECK>
__device__ __forceinline__ void sincos_multi(const double (&x)[K], double (&s)[K], double (&c)[K], bool (&bad)[K]) {
  double big[K], r[K], z[K], ps[K], pc[K];
#pragma unroll
  for (int q = 0; q < K; q++) big[q] = fma(x[q], kSC[0], 6755399441055744.0);
  if (CHECK) {
#pragma unroll
    for (int q = 0; q < K; q++) bad[q] |= !sincos_fast_range(x[q]);
  }
#pragma unroll
  for (int q = 0; q < K; q++) z[q] = big[q] - 6755399441055744.0;             // kd
#pragma unroll
  for (int q = 0; q < K; q++) r[q] = fma(-z[q], kSC[1], x[q]);
#pragma unroll
  for (int q = 0; q < K; q++) r[q] = fma(-z[q], kSC[2], r[q]);
#pragma unroll
  for (int q = 0; q < K; q++) r[q] = fma(-z[q], kSC[3], r[q]);
#pragma unroll
  for (int q = 0; q < K; q++) z[q] = r[q] * r[q];
#pragma unroll
  for (int q = 0; q < K; q++) { ps[q] = fma(z[q], kSC[4], kSC[5]); pc[q] = fma(z[q], kSC[10], kSC[11]); }
#pragma unroll
  for (int q = 0; q < K; q++) { ps[q] = fma(z[q], ps[q], kSC[6]); pc[q] = fma(z[q], pc[q], kSC[12]); }
#pragma unroll
  for (int q = 0; q < K; q++) { ps[q] = fma(z[q], ps[q], kSC[7]); pc[q] = fma(z[q], pc[q], kSC[13]); }
#pragma unroll
  for (int q = 0; q < K; q++) { ps[q] = fma(z[q], ps[q], kSC[8]); pc[q] = fma(z[q], pc[q], kSC[14]); }
#pragma unroll
  for (int q = 0; q < K; q++) { ps[q] = fma(z[q], ps[q], kSC[9]); pc[q] = fma(z[q], pc[q], kSC[15]); }
#pragma unroll
  for (int q = 0; q < K; q++) { s[q] = z[q] * r[q]; pc[q] = fma(z[q], pc[q], -0.5); }
#pragma unroll
  for (int q = 0; q < K; q++) { ps[q] = fma(s[q], ps[q], r[q]); pc[q] = fma(z[q], pc[q], 1.0); }
#pragma unroll
  for (int q = 0; q < K; q++) {
    const int k = __double2loint(big[q]);
    const bool odd = k & 1;
    const double sa = odd ? pc[q] : ps[q], ca = odd ? ps[q] : pc[q];
    s[q] = xor_sign(sa, k << 30);                 // (k & 2) ? -sa : sa
    c[q] = xor_sign(ca, (k + 1) << 30);           // ((k + 1) & 2) ? -ca : ca
  }
}

// np.mod(a, 2*pi) on [-2pi, 4pi) with ONE add: the addend (-2pi / +2pi / 0) is selected, a + 0.0 = a exactly (and
// -0.0 + 0.0 = +0.0 = numpy's copysign(0, b)).  The range check is an integer compare on the high word and conservative
// (a high word equal to the limit's counts as out of range; the slow path is exact, so that only costs time).
__device__ __forceinline__ double mod_two_pi_fast(double a, bool &bad) {
  const bool neg = a < 0.0;
  const double add = (a >= ISLS_TWO_PI) ? -ISLS_TWO_PI : (neg ? ISLS_TWO_PI : 0.0);
  const int ha = __double2hiint(a) & 0x7fffffff;
  bad |= ha >= (neg ? 0x401921fb : 0x402921fb);                  // high words of 2pi and 4pi
  return a + add;
}

struct CarModel {
  static constexpr int n = 4, m = 2, NJ = 6, NJA = 6;
  static constexpr int JX0 = 2;   // jac() reads the state components >= JX0 only
  __host__ __device__ static constexpr int am(int i, int j) {
    return i == j ? MO : ((i <= 1 && j >= 2) || (i == 2 && j == 3)) ? MV : MZ;
  }
  __host__ __device__ static constexpr int bm(int i, int j) {
    return ((i == 2 && j == 0) || (i == 3 && j == 1)) ? MV : MZ;
  }
  // K candidate chains advanced together, stage by stage (see sincos_multi).  Precondition: every chain's heading
  // is inside the fast sincos range - true at entry if fast_state() held for the initial state and afterwards by
  // construction (the heading leaves mod_two_pi_fast in [0, 2pi) or the chain's flag is raised).
  __device__ __forceinline__ static bool fast_state(const double (&x)[n]) { return sincos_fast_range(x[2]); }
  template <int K>
  __device__ __forceinline__ static void steps_fast(double (&x)[K][n], const double (&u)[K][m], double dt,
                                                    bool (&bad)[K]) {
    double th[K], s[K], c[K], dv[K];
#pragma unroll
    for (int q = 0; q < K; q++) th[q] = x[q][2];
    sincos_multi<K, false>(th, s, c, bad);
#pragma unroll
    for (int q = 0; q < K; q++) dv[q] = dt * x[q][3];
#pragma unroll
    for (int q = 0; q < K; q++) th[q] = fma(dv[q], u[q][0], th[q]);
#pragma unroll
    for (int q = 0; q < K; q++) x[q][0] = fma(dv[q], c[q], x[q][0]);
#pragma unroll
    for (int q = 0; q < K; q++) x[q][1] = fma(dv[q], s[q], x[q][1]);
#pragma unroll
    for (int q = 0; q < K; q++) x[q][3] = fma(dt, u[q][1], x[q][3]);
#pragma unroll
    for (int q = 0; q < K; q++) x[q][2] = mod_two_pi_fast(th[q], bad[q]);
  }
  __device__ __forceinline__ static void step(const double (&x)[n], const double (&u)[m], double (&xn)[n],
                                              double dt) {
    double s, c;
    sincos_pio2(x[2], &s, &c);
    const double dv = dt * x[3];
    xn[0] = fma(dv, c, x[0]);
    xn[1] = fma(dv, s, x[1]);
    xn[2] = mod_two_pi(fma(dv, u[0], x[2]));
    xn[3] = fma(dt, u[1], x[3]);
  }
  __device__ __forceinline__ static void jac(const double (&x)[n], const double (&u)[m], double (&J)[NJ],
                                             double dt) {
    double s, c;
    sincos_pio2(x[2], &s, &c);
    const double dv = dt * x[3];
    J[0] = dv * -s;      // A[0][2]
    J[1] = dv * c;       // A[1][2]
    J[2] = dt * c;       // A[0][3]
    J[3] = dt * s;       // A[1][3]
    J[4] = dt * u[0];    // A[2][3]
    J[5] = dv;           // B[2][0]
  }
  __device__ __forceinline__ static void expand(const double (&J)[NJ], double (&A)[n][n], double (&B)[n][m],
                                                double dt) {
    A[0][2] = J[0]; A[1][2] = J[1]; A[0][3] = J[2]; A[1][3] = J[3]; A[2][3] = J[4];
    B[2][0] = J[5]; B[3][1] = dt;
  }
};

// Car-parking model of Tassa et al. as written in the reference's Tutorial (notebooks/Tutorial.ipynb cell 8):
// state [x, y, car angle o, front-wheel velocity v], control [front-wheel angle w, acceleration a], axle distance 2.
//   f = dt v;  S = sqrt(d^2 - (sin w f)^2);  b = f cos w + d - S;  do = asin(sin w f / d)
//   x+ = x + b cos o;  y+ = y + b sin o;  o+ = o + do;  v+ = v + a dt
// The notebook differentiates it with autograd (cell 10); the Jacobian here is the closed form:
//   db/dv = dt (cos w + sin^2 w f / S),  db/dw = -f sin w + sin w cos w f^2 / S,  d(do)/dv = dt sin w / S,
//   d(do)/dw = f cos w / S.
struct TassaCarModel {
  static constexpr int n = 4, m = 2, NJ = 8, NJA = 8;
  static constexpr int JX0 = 2;   // jac() reads the state components >= JX0 only
  static constexpr double DIST = 2.0;
  __host__ __device__ static constexpr int am(int i, int j) {
    return i == j ? MO : ((i <= 1 && j >= 2) || (i == 2 && j == 3)) ? MV : MZ;
  }
  __host__ __device__ static constexpr int bm(int i, int j) {
    return ((i <= 2 && j == 0) || (i == 3 && j == 1)) ? MV : MZ;
  }
  __device__ __forceinline__ static void step(const double (&x)[n], const double (&u)[m], double (&xn)[n],
                                              double dt) {
    double sw, cw, so, co;
    sincos_pio2(u[0], &sw, &cw);
    sincos_pio2(x[2], &so, &co);
    const double f = dt * x[3];
    const double sf = sw * f;
    const double S = sqrt(DIST * DIST - sf * sf);
    const double b = (f * cw + DIST) - S;
    xn[0] = x[0] + b * co;
    xn[1] = x[1] + b * so;
    xn[2] = x[2] + asin(sf / DIST);
    xn[3] = x[3] + u[1] * dt;
  }
  __device__ __forceinline__ static bool fast_state(const double (&)[n]) { return true; }
  template <int K>
  __device__ __forceinline__ static void steps_fast(double (&x)[K][n], const double (&u)[K][m], double dt,
                                                    bool (&)[K]) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double xn[n];
      step(x[k], u[k], xn, dt);
#pragma unroll
      for (int i = 0; i < n; i++) x[k][i] = xn[i];
    }
  }
  __device__ __forceinline__ static void jac(const double (&x)[n], const double (&u)[m], double (&J)[NJ],
                                             double dt) {
    double sw, cw, so, co;
    sincos_pio2(u[0], &sw, &cw);
    sincos_pio2(x[2], &so, &co);
    const double f = dt * x[3];
    const double S = sqrt(DIST * DIST - (sw * f) * (sw * f));
    const double b = (f * cw + DIST) - S;
    const double db_dv = dt * (cw + sw * sw * f / S);
    const double db_dw = -f * sw + sw * cw * f * f / S;
    J[0] = -b * so;          // A[0][2]
    J[1] = b * co;           // A[1][2]
    J[2] = db_dv * co;       // A[0][3]
    J[3] = db_dv * so;       // A[1][3]
    J[4] = dt * sw / S;      // A[2][3]
    J[5] = db_dw * co;       // B[0][0]
    J[6] = db_dw * so;       // B[1][0]
    J[7] = f * cw / S;       // B[2][0]
  }
  __device__ __forceinline__ static void expand(const double (&J)[NJ], double (&A)[n][n], double (&B)[n][m],
                                                double dt) {
    A[0][2] = J[0]; A[1][2] = J[1]; A[0][3] = J[2]; A[1][3] = J[3]; A[2][3] = J[4];
    B[0][0] = J[5]; B[1][0] = J[6]; B[2][0] = J[7]; B[3][1] = dt;
  }
};

struct Arm3Model {
  static constexpr int n = 9, m = 3, NJ = 6, NJA = 6;
  static constexpr int JX0 = 0;   // jac() reads the state components >= JX0 only
  __host__ __device__ static constexpr int am(int i, int j) {
    return i < 6 ? (i == j ? MO : (i < 3 && j == i + 3) ? MV : MZ)
                 : (i < 8 && j < 6) ? MV : MZ;           // rows 6,7 = [J, J dt, 0]; row 8 (p_z) is zero
  }
  __host__ __device__ static constexpr int bm(int i, int j) {
    return i < 3 ? (i == j ? MV : MZ) : i < 6 ? (j == i - 3 ? MV : MZ) : i < 8 ? MV : MZ;
  }
  __device__ __forceinline__ static void qnext(const double (&x)[n], const double (&u)[m], double (&q)[3],
                                               double dt) {
    const double h = dt * dt;
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = fma(0.5 * u[i], h, fma(x[3 + i], dt, x[i]));
  }
  __device__ __forceinline__ static bool fast_state(const double (&)[n]) { return true; }
  template <int K>
  __device__ __forceinline__ static void steps_fast(double (&x)[K][n], const double (&u)[K][m], double dt,
                                                    bool (&bad)[K]) {
    double a[3 * K], s[3 * K], c[3 * K];
    bool b3[3 * K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      double q[3];
      qnext(x[k], u[k], q, dt);
      a[3 * k] = q[0]; a[3 * k + 1] = q[0] + q[1]; a[3 * k + 2] = a[3 * k + 1] + q[2];
#pragma unroll
      for (int i = 0; i < 3; i++) { x[k][i] = q[i]; x[k][3 + i] = fma(u[k][i], dt, x[k][3 + i]); b3[3 * k + i] = false; }
    }
    sincos_multi<3 * K, true>(a, s, c, b3);
#pragma unroll
    for (int k = 0; k < K; k++) {
      x[k][6] = (c[3 * k] + c[3 * k + 1]) + c[3 * k + 2];
      x[k][7] = (s[3 * k] + s[3 * k + 1]) + s[3 * k + 2];
      x[k][8] = 0.0;
      bad[k] |= b3[3 * k] | b3[3 * k + 1] | b3[3 * k + 2];
    }
  }
  __device__ __forceinline__ static void step(const double (&x)[n], const double (&u)[m], double (&xn)[n],
                                              double dt) {
    double q[3];
    qnext(x, u, q, dt);
    const double a1 = q[0], a2 = a1 + q[1], a3 = a2 + q[2];
    double s1, c1, s2, c2, s3, c3;
    sincos_pio2(a1, &s1, &c1); sincos_pio2(a2, &s2, &c2); sincos_pio2(a3, &s3, &c3);
#pragma unroll
    for (int i = 0; i < 3; i++) { xn[i] = q[i]; xn[3 + i] = fma(u[i], dt, x[3 + i]); }
    xn[6] = (c1 + c2) + c3;
    xn[7] = (s1 + s2) + s3;
    xn[8] = 0.0;
  }
  __device__ __forceinline__ static void jac(const double (&x)[n], const double (&u)[m], double (&J)[NJ],
                                             double dt) {
    double q[3];
    qnext(x, u, q, dt);
    const double a1 = q[0], a2 = a1 + q[1], a3 = a2 + q[2];
    double s1, c1, s2, c2, s3, c3;
    sincos_pio2(a1, &s1, &c1); sincos_pio2(a2, &s2, &c2); sincos_pio2(a3, &s3, &c3);
    J[0] = -((s1 + s2) + s3); J[1] = -(s2 + s3); J[2] = -s3;
    J[3] = (c1 + c2) + c3;    J[4] = c2 + c3;    J[5] = c3;
  }
  __device__ __forceinline__ static void expand(const double (&J)[NJ], double (&A)[n][n], double (&B)[n][m],
                                                double dt) {
    const double h = 0.5 * (dt * dt);
#pragma unroll
    for (int i = 0; i < 3; i++) { A[i][i + 3] = dt; B[i][i] = h; B[3 + i][i] = dt; }
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        A[6 + r][c] = J[3 * r + c];
        A[6 + r][3 + c] = J[3 * r + c] * dt;
        B[6 + r][c] = h * J[3 * r + c];        // numpy: 0.5 * J * dt**2 -> (0.5*J)*(dt*dt); scaling by 0.5 is exact
      }
  }
};

template <int D>
struct DoubleIntModel {
  static constexpr int n = 2 * D, m = D, NJ = 0, NJA = 1;
  static constexpr int JX0 = 0;   // jac() reads the state components >= JX0 only
  __host__ __device__ static constexpr int am(int i, int j) { return i == j ? MO : (i < D && j == i + D) ? MV : MZ; }
  __host__ __device__ static constexpr int bm(int i, int j) {
    return (i < D && j == i) ? MV : (i >= D && j == i - D) ? MV : MZ;
  }
  __device__ __forceinline__ static void step(const double (&x)[n], const double (&u)[m], double (&xn)[n],
                                              double dt) {
    const double h = dt * dt / 2.0;               // utils.py:266-276: dt**2 / factorial(2)
#pragma unroll
    for (int i = 0; i < D; i++) {
      xn[i] = (x[i] + dt * x[D + i]) + h * u[i];
      xn[D + i] = x[D + i] + dt * u[i];
    }
  }
  __device__ __forceinline__ static bool fast_state(const double (&)[n]) { return true; }
  template <int K>
  __device__ __forceinline__ static void steps_fast(double (&x)[K][n], const double (&u)[K][m], double dt,
                                                    bool (&)[K]) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double xn[n];
      step(x[k], u[k], xn, dt);
#pragma unroll
      for (int i = 0; i < n; i++) x[k][i] = xn[i];
    }
  }
  __device__ __forceinline__ static void jac(const double (&)[n], const double (&)[m], double (&)[1], double) {}
  __device__ __forceinline__ static void expand(const double (&)[1], double (&A)[n][n], double (&B)[n][m],
                                                double dt) {
    const double h = dt * dt / 2.0;
#pragma unroll
    for (int i = 0; i < D; i++) { A[i][i + D] = dt; B[i][i] = h; B[D + i][i] = dt; }
  }
};

// Generic linear time-invariant model x+ = A x + B u with arbitrary constant A [n, n], B [n, m] (Base.AB of the reference
// takes any pair, isls/base.py:98-119; SLS.ADMM_LQT_DP, isls/sls.py:298-317).  The matrices live in constant memory of
// the translation unit that instantiates the model's kernels (isls_model_lti.cu) and are uploaded, stream-ordered,
// before every launch sequence - one LTI plan per device in flight at a time.
struct LtiConst { double A[36]; double B[18]; };
static __constant__ LtiConst c_lti;
template <int N_, int M_>
struct LtiModel {
  static constexpr int n = N_, m = M_, NJ = 0, NJA = 1;
  static constexpr int JX0 = 0;   // jac() reads the state components >= JX0 only
  __host__ __device__ static constexpr int am(int, int) { return MV; }
  __host__ __device__ static constexpr int bm(int, int) { return MV; }
  __device__ __forceinline__ static void step(const double (&x)[n], const double (&u)[m], double (&xn)[n], double) {
#pragma unroll
    for (int i = 0; i < n; i++) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int j = 0; j < n; j++) a = fma(c_lti.A[i * n + j], x[j], a);
#pragma unroll
      for (int j = 0; j < m; j++) b = fma(c_lti.B[i * m + j], u[j], b);
      xn[i] = a + b;                                   // x.dot(A.T) + u.dot(B.T), isls/sls_base.py:49-53
    }
  }
  __device__ __forceinline__ static bool fast_state(const double (&)[n]) { return true; }
  template <int K>
  __device__ __forceinline__ static void steps_fast(double (&x)[K][n], const double (&u)[K][m], double dt, bool (&)[K]) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double xn[n];
      step(x[k], u[k], xn, dt);
#pragma unroll
      for (int i = 0; i < n; i++) x[k][i] = xn[i];
    }
  }
  __device__ __forceinline__ static void jac(const double (&)[n], const double (&)[m], double (&)[1], double) {}
  __device__ __forceinline__ static void expand(const double (&)[1], double (&A)[n][n], double (&B)[n][m], double) {
#pragma unroll
    for (int i = 0; i < n; i++) {
#pragma unroll
      for (int j = 0; j < n; j++) A[i][j] = c_lti.A[i * n + j];
#pragma unroll
      for (int j = 0; j < m; j++) B[i][j] = c_lti.B[i * m + j];
    }
  }
};
