// lqro_sim.cuh -- simulator-form agent update on the device: the second loop of the
// reference's time step, MAIN:1408-1417, per agent
//     vGoal = newV;  u = findU();  propagateU(u);  kalmanFilter1(x, Rot, u, M, P);
//     z = sampleGaussian(h(xTrue, RotTrue), N);  kalmanFilter2(x, Rot, z, N, P);  vGoal = findVGoal();
// i.e. riccatiControllerSteady (MAIN:547-576), propagate (MAIN:442-455: central-difference Jacobian of the
// quadrotor model f MAIN:338-367, two 16x16 Pade-(7,7) matrix exponentials MAT:773-790, Simpson-weighted process
// noise, sampleGaussian SIM2:19-31 through the cyclic Jacobi eigen-decomposition MAT:674-757), kalmanFilter1/2
// (MAIN:457-487, 6x6 inverse by full-pivot Gauss MAT:603-671) and riccatiControllerSteadyPosition (MAIN:579-608).
//
// Mapping: ONE WARP PER AGENT.  Every 16x16 operand (Jacobian, both exponentials, their Pade workspace, the two
// covariances) lives in the warp's 23 KB of shared memory; a lane owns 8 elements of each 16x16 product (every
// element a left-to-right sum, as Matrix::operator* MAT:219-231 forms it); the 34 evaluations of f behind one
// Jacobian run one (column, side) pair per lane with the 3x3 exponential in registers; pivot searches (full-pivot
// Gauss, Jacobi row pivot) are warp arg-max reductions that keep the reference's first-maximum rule; the rotation
// of one Jacobi step is spread over the 16 lanes of the two rows/columns it touches.  Too small for tensor cores.
//
// ARITHMETIC.  Same contract as the rest of the library (-fmad=false, IEEE / and sqrt, reference operation order),
// so everything except the libm calls (tan, asin, atan2, log, hypot: CUDA's differ from glibc's in the last bit) is
// bit-identical to the CPU reference; the update is held to 1e-9 relative, not bit-exactness (tests/test_gpu_sim.py).
// The Gaussian noise is NOT generated here: the reference draws it from the sequential MSVC rand() stream
// (MAIN:320-336), so the host supplies the standard normals in the reference's order (lqro_msvc_normals) and the
// kernel applies the covariance square root.
//
// DUAL COMPILATION.  This file also compiles as plain C++ (no __CUDACC__): the warp policy collapses to a single
// sequential lane.  Only tests/sim_host uses that build, to check the kernel's logic against the compiled reference
// where no GPU exists; liblqro.so never contains it.
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define SIMF __device__ __forceinline__
#define SIMN __device__ __noinline__
#else
#define SIMF static inline
#define SIMN static
#endif
#if defined(__CUDACC__)
#define SIMM __device__ __forceinline__   // member functions
#else
#define SIMM inline
#endif

// What the reference keeps in globals (MAIN:40-76, setup() MAIN:176-198) and in each Quadrotor (MAIN:95-99).
struct LqroSimModel {
  double dt, gravity, mass, moment_const, thrust_latency, length, jstep;
  double inertia[9], inv_inertia[9];
  double u_goal[4];
  double L[64], E[12], l[4], Lh[48], Eh[9];
  double M[256], N[36];
};

// ---- execution policies ----------------------------------------------------------------------------
struct SimSeq {  // one thread does everything (3x3 work inside a lane; the whole host build)
  SIMM static int lane() { return 0; }
  SIMM static int stride() { return 1; }
  SIMM static void sync() {}
  SIMM static void argmax(double&, int&) {}
};
#if defined(__CUDACC__)
struct SimWarp {
  SIMM static int lane() { return (int)(threadIdx.x & 31); }
  SIMM static int stride() { return 32; }
  SIMM static void sync() { __syncwarp(); }
  // larger value wins, equal values go to the smaller index: the reference's scans use strict > in
  // increasing index order, i.e. the first maximum
  SIMM static void argmax(double& v, int& idx) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
  }
};
SIMF double sim_hypot(double x, double y) {  // hypot(x, y) with an FMA-evaluated residual correction (MAT:715,717 call the CRT's)
  x = fabs(x); y = fabs(y);
  if (x < y) { const double t = x; x = y; y = t; }
  if (x == 0.0 || isinf(x)) return x;
  const double h = sqrt(fma(x, x, y * y));
  const double h2 = h * h, x2 = x * x;
  const double r = fma(-y, y, h2 - x2) + fma(h, h, -h2) - fma(x, x, -x2);
  return h - r / (2.0 * h);
}
#else
typedef SimSeq SimWarp;
SIMF double sim_hypot(double x, double y) { return ::hypot(x, y); }
#endif

#define SIM_FOR(P, e, n) for (int e = P::lane(); e < (n); e += P::stride())

// ---- small dense algebra in the reference's operation order ---------------------------------------
// C (n x p) = A (n x m) * B (m x p); TB: B is given transposed (stored p x m).  MAT:219-231.
template <class P, bool TB>
SIMF void sim_mm(double* C, const double* A, const double* B, int n, int m, int p) {
  SIM_FOR(P, e, n * p) {
    const int i = e / p, j = e - i * p;
    double t = 0.0;
    for (int k = 0; k < m; ++k) t += A[i * m + k] * (TB ? B[j * m + k] : B[k * p + j]);
    C[e] = t;
  }
  P::sync();
}

// P X = Q by Gaussian elimination with FULL pivoting, the algorithm shared by operator% (MAT:370-440) and
// operator! (MAT:603-671; with Q = I the two perform the same operations up to the sign of zeros).
// m: in P, out X;  inv: in Q, destroyed.  rp/cp: N ints each, fac: N doubles of scratch.
template <class P, int N>
SIMF void sim_solve(double* m, double* inv, int* rp, int* cp, double* fac) {
  SIM_FOR(P, i, N) { rp[i] = i; cp[i] = i; }
  P::sync();
  for (int k = 0; k < N; ++k) {
    const int w = N - k;
    double best = 0.0; int bt = 0x7fffffff;
    SIM_FOR(P, t, w * w) {  // (i, j) in row-major order of the trailing block, first maximum
      const int i = k + t / w, j = k + t % w;
      const double a = fabs(m[rp[i] * N + cp[j]]);
      if (a > best) { best = a; bt = t; }
    }
    P::argmax(best, bt);
    if (!(best > 0.0)) bt = 0;
    P::sync();
    if (P::lane() == 0) {
      const int mr = k + bt / w, mc = k + bt % w;
      int s = rp[k]; rp[k] = rp[mr]; rp[mr] = s;
      s = cp[k]; cp[k] = cp[mc]; cp[mc] = s;
    }
    P::sync();
    const int rk = rp[k], ck = cp[k];
    SIM_FOR(P, i, N) { if (i > k) fac[i] = m[rp[i] * N + ck] / m[rk * N + ck]; }
    P::sync();
    SIM_FOR(P, t, (w - 1) * (w - 1 + N)) {
      const int i = k + 1 + t / (w - 1 + N), jj = t % (w - 1 + N);
      const double f = fac[i];
      if (jj < w - 1) { const int c = cp[k + 1 + jj]; m[rp[i] * N + c] -= f * m[rk * N + c]; }
      else { const int c = jj - (w - 1); inv[rp[i] * N + c] -= f * inv[rk * N + c]; }
    }
    P::sync();
  }
  for (int k = N - 1; k >= 0; --k) {
    const int rk = rp[k], ck = cp[k];
    const double quotient = m[rk * N + ck];
    SIM_FOR(P, j, N) inv[rk * N + j] /= quotient;
    P::sync();
    SIM_FOR(P, t, k * N) {
      const int i = t / N, j = t - i * N;
      inv[rp[i] * N + j] -= m[rp[i] * N + ck] * inv[rk * N + j];
    }
    P::sync();
  }
  SIM_FOR(P, t, N * N) { const int i = t / N, j = t - i * N; m[cp[i] * N + j] = inv[rp[i] * N + j]; }
  P::sync();
}

#define SIM_JB0 1729728e1
#define SIM_JB1 864864e1
#define SIM_JB2 199584e1
#define SIM_JB3 2772e2
#define SIM_JB4 252e2
#define SIM_JB5 1512e0
#define SIM_JB6 56e0
#define SIM_JB7 1e0
#define SIM_NORMLIM 9.504178996162932e-1
#define SIM_LOG2E 1.4426950408889634074

// out = exp(q) by scaling, the (7,7) Pade approximant and squaring (MAT:773-790).  ws: 6*N*N doubles, ints 2N,
// small N doubles.  out may not alias q or ws.
template <class P, int N>
SIMF void sim_exp(double* out, const double* q, double* ws, int* perm, double* small) {
  double* A = ws; double* A2 = ws + N * N; double* A4 = ws + 2 * N * N; double* A6 = ws + 3 * N * N;
  double* T = ws + 4 * N * N; double* U = ws + 5 * N * N;
  SIM_FOR(P, j, N) {  // 1-norm: column sums top to bottom (MAT:275-288)
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += fabs(q[i * N + j]);
    small[j] = s;
  }
  P::sync();
  double norm1 = 0.0;
  for (int j = 0; j < N; ++j) if (small[j] > norm1) norm1 = small[j];
  const double sd = ceil(log(norm1 / SIM_NORMLIM) * SIM_LOG2E);
  const int s = (int)(sd > 0.0 ? sd : 0.0);
  const double scale = ldexp(1.0, s);  // pow(2.0, s), exact
  SIM_FOR(P, e, N * N) A[e] = q[e] / scale;
  P::sync();
  sim_mm<P, false>(A2, A, A, N, N, N);
  sim_mm<P, false>(A4, A2, A2, N, N, N);
  sim_mm<P, false>(A6, A2, A4, N, N, N);
  SIM_FOR(P, e, N * N) {
    const double id = (e / N == e % N) ? 1.0 : 0.0;
    T[e] = A6[e] * SIM_JB7 + A4[e] * SIM_JB5 + A2[e] * SIM_JB3 + id * SIM_JB1;
  }
  P::sync();
  sim_mm<P, false>(U, A, T, N, N, N);
  SIM_FOR(P, e, N * N) {  // V, then (V - U) into A2 and (V + U) into A4 (A2, A4 are dead once V is formed)
    const double id = (e / N == e % N) ? 1.0 : 0.0;
    const double v = A6[e] * SIM_JB6 + A4[e] * SIM_JB4 + A2[e] * SIM_JB2 + id * SIM_JB0;
    T[e] = v;
  }
  P::sync();
  SIM_FOR(P, e, N * N) { A2[e] = T[e] - U[e]; A4[e] = T[e] + U[e]; }
  P::sync();
  sim_solve<P, N>(A2, A4, perm, perm + N, small);  // R7 = (V - U) % (V + U), result in A2
  double* cur = A2; double* nxt = A6;
  for (int i = 0; i < s; ++i) {
    sim_mm<P, false>(nxt, cur, cur, N, N, N);
    double* t = cur; cur = nxt; nxt = t;
  }
  SIM_FOR(P, e, N * N) out[e] = cur[e];
  P::sync();
}

// Cyclic-by-row Jacobi eigen-decomposition (MAT:674-757): D <- eigenvalues on the diagonal (upper triangle
// worked on), V <- eigenvectors.  D holds the matrix on entry.
template <class P, int N>
SIMF void sim_jacobi(double* D, double* V) {
  SIM_FOR(P, e, N * N) V[e] = (e / N == e % N) ? 1.0 : 0.0;
  P::sync();
  int pivotRow = 0, zeroCount = 0;
  while (true) {
    double maximum = 0.0; int bi = 0x7fffffff;
    SIM_FOR(P, k, N) {
      if (k == pivotRow) continue;
      const double a = fabs(k < pivotRow ? D[k * N + pivotRow] : D[pivotRow * N + k]);
      if (a > maximum) { maximum = a; bi = k; }
    }
    P::argmax(maximum, bi);
    const int p = bi < pivotRow ? bi : pivotRow, q = bi < pivotRow ? pivotRow : bi;
    pivotRow = (pivotRow + 1) % N;
    if (maximum <= DBL_EPSILON) {
      if (++zeroCount == N) break;
      continue;
    }
    zeroCount = 0;
    const double Dpq = D[p * N + q];
    const double theta = 0.5 * (D[q * N + q] - D[p * N + p]) / Dpq;
    double t = 1 / (fabs(theta) + sim_hypot(theta, 1));
    if (theta < 0) t = -t;
    const double c = 1 / sim_hypot(t, 1);
    const double s = c * t;
    const double tau = s / (1 + c);
    P::sync();
    SIM_FOR(P, r, 2 * N) {
      if (r < N) {
        if (r < p) {
          const double Drp = D[r * N + p], Drq = D[r * N + q];
          D[r * N + p] -= s * (Drq + tau * Drp);
          D[r * N + q] += s * (Drp - tau * Drq);
        } else if (r > p && r < q) {
          const double Drp = D[p * N + r], Drq = D[r * N + q];
          D[p * N + r] -= s * (Drq + tau * Drp);
          D[r * N + q] += s * (Drp - tau * Drq);
        } else if (r > q) {
          const double Drp = D[p * N + r], Drq = D[q * N + r];
          D[p * N + r] -= s * (Drq + tau * Drp);
          D[q * N + r] += s * (Drp - tau * Drq);
        } else if (r == p) {  // the lane that owns no off-diagonal pair updates the 2x2 pivot block
          D[p * N + p] -= t * Dpq;
          D[q * N + q] += t * Dpq;
          D[p * N + q] = 0;
        }
      } else {
        const int rr = r - N;
        const double Vrp = V[rr * N + p], Vrq = V[rr * N + q];
        V[rr * N + p] -= s * (Vrq + tau * Vrp);
        V[rr * N + q] += s * (Vrp - tau * Vrq);
      }
    }
    P::sync();
  }
}

// sample = SVec * sqrt(SVal) * normals + mean (SIM2:19-31); var is destroyed (it becomes D), V: N*N scratch
template <class P, int N>
SIMF void sim_sample_gaussian(double* out, const double* mean, double* var, double* V, const double* normals) {
  sim_jacobi<P, N>(var, V);
  SIM_FOR(P, i, N) {
    double t = 0.0;
    for (int j = 0; j < N; ++j) t += (V[i * N + j] * sqrt(var[j * N + j])) * normals[j];
    out[i] = t + mean[i];
  }
  P::sync();
}

// ---- per-thread 3x3 helpers -----------------------------------------------------------------------------
SIMF void sim_skew(double* S, const double* v) {  // stdafx.h:77-84
  S[0] = 0.0; S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2]; S[4] = 0.0; S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0]; S[8] = 0.0;
}
SIMF void sim_mv3(double* y, const double* Mx, const double* v) {
  for (int i = 0; i < 3; ++i) { double t = 0.0; for (int k = 0; k < 3; ++k) t += Mx[i * 3 + k] * v[k]; y[i] = t; }
}
SIMF void sim_mtv3(double* y, const double* Mx, const double* v) {  // (~M) * v
  for (int i = 0; i < 3; ++i) { double t = 0.0; for (int k = 0; k < 3; ++k) t += Mx[k * 3 + i] * v[k]; y[i] = t; }
}
SIMF double sim_hypot3(const double* v) {  // stdafx.h:86-97
  const double X = fabs(v[0]), Y = fabs(v[1]), Z = fabs(v[2]);
  if (X > Y && X > Z) return X * sqrt(1.0 + (Y / X) * (Y / X) + (Z / X) * (Z / X));
  else if (Y > Z) return Y * sqrt(1.0 + (X / Y) * (X / Y) + (Z / Y) * (Z / Y));
  else return Z * sqrt(1.0 + (X / Z) * (X / Z) + (Y / Z) * (Y / Z));
}
SIMF void sim_exp3(double* out, const double* q) {
  double ws[54]; int perm[6]; double small[3];
  sim_exp<SimSeq, 3>(out, q, ws, perm, small);
}
// R <- R * exp(skew(r))   (the "reset rotation-error" step, MAIN:453, 472, 485)
SIMF void sim_fold_rotation(double* R, const double* r) {
  double S[9], Ex[9], Rn[9];
  sim_skew(S, r);
  sim_exp3(Ex, S);
  sim_mm<SimSeq, false>(Rn, R, Ex, 3, 3, 3);
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
}

// Quadrotor model xdot = f(x, R, u)  (MAIN:338-367)
SIMF void sim_f(double* xdot, const double* x, const double* R, const double* u, const LqroSimModel& md) {
  const double* v = x + 3; const double* r = x + 6; const double* w = x + 9; const double* F = x + 12;
  xdot[0] = v[0]; xdot[1] = v[1]; xdot[2] = v[2];
  {  // -gravity*eZ + R*exp([r])*((F0+F1+F2+F3)/mass)*eZ
    double S[9], Ex[9], RE[9];
    sim_skew(S, r);
    sim_exp3(Ex, S);
    sim_mm<SimSeq, false>(RE, R, Ex, 3, 3, 3);
    const double a = (F[0] + F[1] + F[2] + F[3]) / md.mass;
    const double ez[3] = {0.0, 0.0, 1.0};
    for (int i = 0; i < 3; ++i) {
      double t = 0.0;
      for (int k = 0; k < 3; ++k) t += (RE[i * 3 + k] * a) * ez[k];
      xdot[3 + i] = ez[i] * (-md.gravity) + t;
    }
  }
  {
    const double l = sim_hypot3(r);
    double S[9], Sh[9], t1[3];
    sim_skew(S, r);
    for (int i = 0; i < 9; ++i) Sh[i] = S[i] * 0.5;
    sim_mv3(t1, Sh, w);
    if (0.5 * l > 0.0) {
      double rl[3] = {r[0] / l, r[1] / l, r[2] / l}, Sl[9], Sc[9], t2[3], t3[3];
      sim_skew(Sl, rl);
      const double c = 1.0 - 0.5 * l / tan(0.5 * l);
      for (int i = 0; i < 9; ++i) Sc[i] = Sl[i] * c;
      sim_mv3(t2, Sl, w);
      sim_mv3(t3, Sc, t2);
      for (int i = 0; i < 3; ++i) xdot[6 + i] = w[i] + t1[i] + t3[i];
    } else {
      for (int i = 0; i < 3; ++i) xdot[6 + i] = w[i] + t1[i];
    }
  }
  {  // invInertia*( length*(F1-F3)*eX + length*(F2-F0)*eY + (F0-F1+F2-F3)*momentConst*eZ - [w]*inertia*w )
    const double a = md.length * (F[1] - F[3]), b = md.length * (F[2] - F[0]), c = (F[0] - F[1] + F[2] - F[3]) * md.moment_const;
    const double ex[3] = {1.0, 0.0, 0.0}, ey[3] = {0.0, 1.0, 0.0}, ez[3] = {0.0, 0.0, 1.0};
    double S[9], SJ[9], g[3], rhs[3];
    sim_skew(S, w);
    sim_mm<SimSeq, false>(SJ, S, md.inertia, 3, 3, 3);
    sim_mv3(g, SJ, w);
    for (int i = 0; i < 3; ++i) rhs[i] = ex[i] * a + ey[i] * b + ez[i] * c - g[i];
    sim_mv3(xdot + 9, md.inv_inertia, rhs);
  }
  for (int i = 0; i < 4; ++i) xdot[12 + i] = (u[i] - F[i]) * md.thrust_latency;
}

// The yaw-free local frame both controllers build (MAIN:550-563, 582-595) and the local error state
SIMF void sim_local_frame(double* RLocal, double* xTilde, const double* x, const double* R0, const LqroSimModel& md) {
  const double zR0[3] = {R0[2], R0[5], R0[8]}, zVec[3] = {0.0, 0.0, 1.0};
  double S[9], axis[3];
  sim_skew(S, zR0);
  sim_mv3(axis, S, zVec);
  const double sinangle = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  const double angle = asin(sinangle);
  if (sinangle != 0) { const double f = angle / sinangle; axis[0] = axis[0] * f; axis[1] = axis[1] * f; axis[2] = axis[2] * f; }
  double Ex[9];
  sim_skew(S, axis);
  sim_exp3(Ex, S);
  sim_mm<SimSeq, false>(RLocal, Ex, R0, 3, 3, 3);
  sim_mtv3(xTilde, RLocal, x);
  sim_mtv3(xTilde + 3, RLocal, x + 3);
  {  // errFromRot(~RLocal*R0)  (stdafx.h:35-49)
    double Rr[9], q[3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { double t = 0.0; for (int k = 0; k < 3; ++k) t += RLocal[k * 3 + i] * R0[k * 3 + j]; Rr[i * 3 + j] = t; }
    q[0] = Rr[7] - Rr[5]; q[1] = Rr[2] - Rr[6]; q[2] = Rr[3] - Rr[1];
    const double rr = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double tt = Rr[0] + Rr[4] + Rr[8] - 1;
    if (rr == 0) { xTilde[6] = 0.0; xTilde[7] = 0.0; xTilde[8] = 0.0; }
    else { const double f = atan2(rr, tt) / rr; xTilde[6] = q[0] * f; xTilde[7] = q[1] * f; xTilde[8] = q[2] * f; }
  }
  for (int i = 0; i < 3; ++i) xTilde[9 + i] = x[9 + i];
  for (int i = 0; i < 4; ++i) xTilde[12 + i] = x[12 + i] - md.u_goal[i];
}
// u = uGoal + L*xTilde + E*vTildeGoal + l   (riccatiControllerSteady, MAIN:547-576)
SIMF void sim_find_u(double* u, const double* x, const double* R0, const double* vGoal, const LqroSimModel& md) {
  double RLocal[9], xTilde[16], vT[3];
  sim_local_frame(RLocal, xTilde, x, R0, md);
  sim_mtv3(vT, RLocal, vGoal);
  for (int i = 0; i < 4; ++i) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 16; ++k) a += md.L[i * 16 + k] * xTilde[k];
    for (int k = 0; k < 3; ++k) b += md.E[i * 3 + k] * vT[k];
    u[i] = md.u_goal[i] + a + b + md.l[i];
  }
}
// vGoal = RLocal*(Lh*xTilde + Eh*pTildeGoal)   (riccatiControllerSteadyPosition, MAIN:579-608)
SIMF void sim_find_vgoal(double* vG, const double* x, const double* R0, const double* pGoal, const LqroSimModel& md) {
  double RLocal[9], xTilde[16], pT[3], s[3];
  sim_local_frame(RLocal, xTilde, x, R0, md);
  sim_mtv3(pT, RLocal, pGoal);
  for (int i = 0; i < 3; ++i) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 16; ++k) a += md.Lh[i * 16 + k] * xTilde[k];
    for (int k = 0; k < 3; ++k) b += md.Eh[i * 3 + k] * pT[k];
    s[i] = a + b;
  }
  sim_mv3(vG, RLocal, s);
}

// ---- per-agent (per-warp) workspace -------------------------------------------------------------------
struct LqroSimWs {
  double F[256], Ae[256], A2e[256], MM[256], P[256], W[6 * 256];
  double x[16], xT[16], R[9], RT[9], u[4], z[6], hz[6], nrm[22], xdot[16], v1[16], v2[16], samp[16], small[16];
  int perm[32];
};
// 6-dimensional operands of the observation / kalmanFilter2 overlay the part of W the 16x16 products there leave free
struct LqroSimObs {
  double *H, *PHt, *HP, *S6, *S6b, *K, *V6;
  SIMM explicit LqroSimObs(double* W) : H(W + 512), PHt(W + 608), HP(W + 704), S6(W + 800), S6b(W + 836), K(W + 872), V6(W + 968) {}
};

// Common half of propagate and kalmanFilter1 (MAIN:443-450, 458-468): F = Jacobian_fx, xDot = f, A = exp(dt F),
// A2 = exp(dt/2 F), MM = (dt/6)(M + 4 A2 M ~A2 + A M ~A), v1 = (dt/6)(xDot + 4 (A2 xDot) + A xDot).
// xs, Rs: the state / rotation linearised about (shared memory of the warp).
template <class P>
SIMN void sim_flow(LqroSimWs& w, const double* xs, const double* Rs, const LqroSimModel& md) {
  SIM_FOR(P, c, 17) {  // columns 0..15 of the central-difference Jacobian (MAIN:392-401), "column" 16 = f itself
    double xr[16], R[9], u[4], fr[16], fl[16];
    for (int i = 0; i < 9; ++i) R[i] = Rs[i];
    for (int i = 0; i < 4; ++i) u[i] = w.u[i];
    for (int side = 0; side < 2; ++side) {
      for (int i = 0; i < 16; ++i) xr[i] = xs[i];
      if (c < 16) {
        for (int i = 0; i < 16; ++i) if (i == c) { if (side == 0) xr[i] += md.jstep; else xr[i] -= md.jstep; }
      }
      sim_f(side == 0 ? fr : fl, xr, R, u, md);
      if (c == 16) break;
    }
    if (c < 16) { for (int i = 0; i < 16; ++i) w.F[i * 16 + c] = (fr[i] - fl[i]) / (2 * md.jstep); }
    else { for (int i = 0; i < 16; ++i) w.xdot[i] = fr[i]; }
  }
  P::sync();
  double* T1 = w.W; double* T2 = w.W + 256;  // sim_exp uses all of W: scaled copies go to MM / P-free buffers instead
  SIM_FOR(P, e, 256) w.MM[e] = w.F[e] * md.dt;
  P::sync();
  sim_exp<P, 16>(w.Ae, w.MM, w.W, w.perm, w.small);
  SIM_FOR(P, e, 256) w.MM[e] = w.F[e] * (md.dt * 0.5);
  P::sync();
  sim_exp<P, 16>(w.A2e, w.MM, w.W, w.perm, w.small);
  // MM = (dt/6) * (M + 4*A2*M*~A2 + A*M*~A)
  double* T3 = w.W + 512; double* T4 = w.W + 768;
  SIM_FOR(P, e, 256) T1[e] = w.A2e[e] * 4;
  P::sync();
  sim_mm<P, false>(T2, T1, md.M, 16, 16, 16);
  sim_mm<P, true>(T3, T2, w.A2e, 16, 16, 16);
  sim_mm<P, false>(T2, w.Ae, md.M, 16, 16, 16);
  sim_mm<P, true>(T4, T2, w.Ae, 16, 16, 16);
  SIM_FOR(P, e, 256) w.MM[e] = (md.M[e] + T3[e] + T4[e]) * (md.dt / 6);
  // v1 = (dt/6)*(xDot + 4*(A2*xDot) + A*xDot)
  SIM_FOR(P, i, 16) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 16; ++k) a += w.A2e[i * 16 + k] * w.xdot[k];
    for (int k = 0; k < 16; ++k) b += w.Ae[i * 16 + k] * w.xdot[k];
    w.v1[i] = (w.xdot[i] + a * 4 + b) * (md.dt / 6);
  }
  P::sync();
}

// One agent, MAIN:1409-1416.  On entry the workspace holds x, R (estimate), xT, RT (true state), P, nrm (22 standard
// normals: 16 for propagate's process noise, 6 for the observation); newV/pGoal: 3 doubles each.
// On exit x, R, xT, RT, P are updated, w.u holds the control input, vG the next goal velocity.
template <class P>
SIMF void sim_update_agent(LqroSimWs& w, const double* newV, const double* pGoal, double* vG, const LqroSimModel& md) {
  {  // Input u = findU()  (vGoal = newV, MAIN:1409-1410)
    double x[16], R[9], u[4];
    for (int i = 0; i < 16; ++i) x[i] = w.x[i];
    for (int i = 0; i < 9; ++i) R[i] = w.R[i];
    sim_find_u(u, x, R, newV, md);
    P::sync();
    if (P::lane() == 0) for (int i = 0; i < 4; ++i) w.u[i] = u[i];
    P::sync();
  }
  // propagate(xTrue, RotTrue, u, M)  (MAIN:442-455)
  sim_flow<P>(w, w.xT, w.RT, md);
  {
    double* T1 = w.W; double* V = w.W + 256;
    SIM_FOR(P, e, 256) T1[e] = w.MM[e];
    SIM_FOR(P, i, 16) w.v2[i] = 0.0;
    P::sync();
    sim_sample_gaussian<P, 16>(w.samp, w.v2, T1, V, w.nrm);
    double r[3], R[9];
    for (int i = 0; i < 3; ++i) r[i] = w.xT[6 + i] + w.v1[6 + i] + w.samp[6 + i];
    for (int i = 0; i < 9; ++i) R[i] = w.RT[i];
    sim_fold_rotation(R, r);
    P::sync();
    SIM_FOR(P, i, 16) w.xT[i] = (i >= 6 && i < 9) ? 0.0 : w.xT[i] + w.v1[i] + w.samp[i];
    SIM_FOR(P, i, 9) w.RT[i] = R[i];
    P::sync();
  }
  // kalmanFilter1(x, Rot, u, M, P)  (MAIN:457-474)
  sim_flow<P>(w, w.x, w.R, md);
  {
    double* T1 = w.W; double* T2 = w.W + 256;
    sim_mm<P, false>(T1, w.Ae, w.P, 16, 16, 16);
    sim_mm<P, true>(T2, T1, w.Ae, 16, 16, 16);
    SIM_FOR(P, e, 256) w.P[e] = T2[e] + w.MM[e];
    double r[3], R[9];
    for (int i = 0; i < 3; ++i) r[i] = w.x[6 + i] + w.v1[6 + i];
    for (int i = 0; i < 9; ++i) R[i] = w.R[i];
    sim_fold_rotation(R, r);
    P::sync();
    SIM_FOR(P, i, 16) w.x[i] = (i >= 6 && i < 9) ? 0.0 : w.x[i] + w.v1[i];
    SIM_FOR(P, i, 9) w.R[i] = R[i];
    P::sync();
  }
  const LqroSimObs o(w.W);
  // Observation z = sampleGaussian(h(xTrue, RotTrue), N)  (MAIN:1413; h MAIN:369-390: rate gyros, then position)
  {
    SIM_FOR(P, i, 6) w.hz[i] = i < 3 ? w.xT[9 + i] : w.xT[i - 3];
    SIM_FOR(P, e, 36) o.S6[e] = md.N[e];
    P::sync();
    sim_sample_gaussian<P, 6>(w.z, w.hz, o.S6, o.V6, w.nrm + 16);
  }
  // kalmanFilter2(x, Rot, z, N, P)  (MAIN:476-487)
  {
    SIM_FOR(P, e, 96) {  // H = Jacobian_hx: central differences of h (MAIN:414-423)
      const int i = e / 16, c = e - i * 16;
      const int src = i < 3 ? 9 + i : i - 3;
      const double xc = w.x[src];
      const double hr = (src == c) ? xc + md.jstep : xc, hl = (src == c) ? xc - md.jstep : xc;
      o.H[e] = (hr - hl) / (2 * md.jstep);
    }
    P::sync();
    sim_mm<P, true>(o.PHt, w.P, o.H, 16, 16, 6);   // P*~H      16x6
    sim_mm<P, false>(o.HP, o.H, w.P, 6, 16, 16);   // H*P       6x16
    sim_mm<P, true>(o.S6, o.HP, o.H, 6, 16, 6);    // (H*P)*~H  6x6
    SIM_FOR(P, e, 36) { o.S6[e] = o.S6[e] + md.N[e]; o.S6b[e] = (e / 6 == e % 6) ? 1.0 : 0.0; }
    P::sync();
    sim_solve<P, 6>(o.S6, o.S6b, w.perm, w.perm + 6, w.small);  // !(H*P*~H + N), result in S6
    sim_mm<P, false>(o.K, o.PHt, o.S6, 16, 6, 6);  // K 16x6
    SIM_FOR(P, i, 6) w.v2[i] = w.z[i] - (i < 3 ? w.x[9 + i] : w.x[i - 3]);
    P::sync();
    SIM_FOR(P, i, 16) { double t = 0.0; for (int k = 0; k < 6; ++k) t += o.K[i * 6 + k] * w.v2[k]; w.v1[i] = t; }
    double* T1 = w.W; double* T2 = w.W + 256;
    sim_mm<P, false>(T1, o.K, o.H, 16, 6, 16);     // K*H
    SIM_FOR(P, e, 256) T1[e] = ((e / 16 == e % 16) ? 1.0 : 0.0) - T1[e];
    P::sync();
    sim_mm<P, false>(T2, T1, w.P, 16, 16, 16);
    SIM_FOR(P, e, 256) w.P[e] = T2[e];
    double r[3], R[9];
    for (int i = 0; i < 3; ++i) r[i] = w.x[6 + i] + w.v1[6 + i];
    for (int i = 0; i < 9; ++i) R[i] = w.R[i];
    sim_fold_rotation(R, r);
    P::sync();
    SIM_FOR(P, i, 16) w.x[i] = (i >= 6 && i < 9) ? 0.0 : w.x[i] + w.v1[i];
    SIM_FOR(P, i, 9) w.R[i] = R[i];
    P::sync();
  }
  {  // vGoal = findVGoal()  (MAIN:1416)
    double x[16], R[9];
    for (int i = 0; i < 16; ++i) x[i] = w.x[i];
    for (int i = 0; i < 9; ++i) R[i] = w.R[i];
    sim_find_vgoal(vG, x, R, pGoal, md);
  }
}

#if defined(__CUDACC__)
// grid-stride over agents, one warp each; dynamic shared memory = sizeof(LqroSimModel) + warps * sizeof(LqroSimWs)
__global__ void __launch_bounds__(128) k_sim_update(int n, const LqroSimModel* __restrict__ model, double* __restrict__ x,
                                                    double* __restrict__ Rot, double* __restrict__ Pcov, double* __restrict__ xTrue,
                                                    double* __restrict__ RotTrue, const double* __restrict__ newV,
                                                    const double* __restrict__ pGoal, const double* __restrict__ normals,
                                                    double* __restrict__ vGoal, double* __restrict__ uOut) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  LqroSimModel& md = *reinterpret_cast<LqroSimModel*>(sim_smem);
  {
    const double* src = reinterpret_cast<const double*>(model);
    double* dst = reinterpret_cast<double*>(sim_smem);
    for (int e = threadIdx.x; e < (int)(sizeof(LqroSimModel) / 8); e += blockDim.x) dst[e] = src[e];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  LqroSimWs& w = *reinterpret_cast<LqroSimWs*>(sim_smem + sizeof(LqroSimModel) + (size_t)warp * sizeof(LqroSimWs));
  for (int a = blockIdx.x * wpb + warp; a < n; a += gridDim.x * wpb) {
    for (int e = lane; e < 256; e += 32) w.P[e] = Pcov[(size_t)a * 256 + e];
    if (lane < 16) { w.x[lane] = x[(size_t)a * 16 + lane]; w.xT[lane] = xTrue[(size_t)a * 16 + lane]; }
    if (lane < 9) { w.R[lane] = Rot[(size_t)a * 9 + lane]; w.RT[lane] = RotTrue[(size_t)a * 9 + lane]; }
    if (lane < 22) w.nrm[lane] = normals[(size_t)a * 22 + lane];
    __syncwarp();
    double nv[3], pg[3], vG[3];
    for (int i = 0; i < 3; ++i) { nv[i] = newV[(size_t)a * 3 + i]; pg[i] = pGoal[(size_t)a * 3 + i]; }
    sim_update_agent<SimWarp>(w, nv, pg, vG, md);
    for (int e = lane; e < 256; e += 32) Pcov[(size_t)a * 256 + e] = w.P[e];
    if (lane < 16) { x[(size_t)a * 16 + lane] = w.x[lane]; xTrue[(size_t)a * 16 + lane] = w.xT[lane]; }
    if (lane < 9) { Rot[(size_t)a * 9 + lane] = w.R[lane]; RotTrue[(size_t)a * 9 + lane] = w.RT[lane]; }
    if (lane < 3) vGoal[(size_t)a * 3 + lane] = vG[lane];
    if (uOut && lane < 4) uOut[(size_t)a * 4 + lane] = w.u[lane];
    __syncwarp();
  }
}
#endif
