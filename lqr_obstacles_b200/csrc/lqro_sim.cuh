// lqro_sim.cuh -- simulator-form agent update on the device: the second loop of the
// reference's time step, MAIN:1408-1417, per agent
//     vGoal = newV;  u = findU();  propagateU(u);  kalmanFilter1(x, Rot, u, M, P);
//     z = sampleGaussian(h(xTrue, RotTrue), N);  kalmanFilter2(x, Rot, z, N, P);  vGoal = findVGoal();
// i.e. riccatiControllerSteady (MAIN:547-576), propagate (MAIN:442-455: central-difference Jacobian of the
// quadrotor model f MAIN:338-367, two 16x16 Pade-(7,7) matrix exponentials MAT:773-790, Simpson-weighted process
// noise, sampleGaussian SIM2:19-31 through the cyclic Jacobi eigen-decomposition MAT:674-757), kalmanFilter1/2
// (MAIN:457-487, 6x6 inverse by full-pivot Gauss MAT:603-671) and riccatiControllerSteadyPosition (MAIN:579-608).
//
// Mapping: FOUR KERNELS, each sized by the shared memory its phase needs (the phases are latency-bound chains of small FP64
// operations -- full-pivot Gauss, Jacobi rotations, 16-term sums -- so resident warps per SM decide the throughput):
//   k_sim_control  thread per agent   u = findU()
//   k_sim_flow     warp per TASK      the linearised flow about the true state (propagate) and about the estimate
//                                     (kalmanFilter1) are independent: 2 tasks per agent.  One flow = 32 evaluations of f
//                                     (one per lane, 3x3 exponential in registers), exp(dt F) and exp(dt/2 F) from ONE 16x16
//                                     Pade approximant + solve (sim_exp_with_half: exact work avoidance), the noise
//                                     covariance MM and the Simpson increment; 5 16x16 buffers = 11 KB per warp, 20 warps/SM
//   k_sim_noise    8 lanes per agent  Jacobi eigen-decomposition of propagate's MM + the sample: its scalar division /
//                                     square-root chain is executed once per warp for four agents in lockstep
//   k_sim_filter   warp per agent     covariance / state updates of kalmanFilter1 and propagate, the observation sample,
//                                     kalmanFilter2 (6x6 full-pivot inverse), findVGoal; 11 KB per warp, 20 warps/SM
// A lane owns 8 elements of each 16x16 product (every element a left-to-right sum, as Matrix::operator* MAT:219-231
// forms it); pivot searches (full-pivot Gauss, Jacobi row pivot) are REDUX arg-max reductions that keep the reference's
// first-maximum rule.  Too small for tensor cores.
//
// ARITHMETIC.  Same contract as the rest of the library (-fmad=false, IEEE / and sqrt, reference operation order),
// so everything except the libm calls (tan, asin, atan2, log, hypot: CUDA's differ from glibc's in the last bit) is
// bit-identical to the CPU reference; the update is held to 1e-9 relative, not bit-exactness (tests/test_gpu_sim.py).
// The Gaussian noise is NOT generated here: the reference draws it from the sequential MSVC rand() stream
// (MAIN:320-336), so the host supplies the standard normals in the reference's order (lqro_msvc_normals) and the
// kernel applies the covariance square root.
//
// DUAL COMPILATION.  This file also compiles as plain C++ (no __CUDACC__): the warp policy collapses to a single
// sequential lane (the warp-specialised routines -- sim_solve_warp, sim_mm16_warp, k_sim_noise's packed Jacobi -- have
// generic twins that the host build uses).  Only tests/sim_host uses that build, to check the kernel's logic against the compiled reference
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

// Per-phase SM-cycle counters of the update kernel (lane 0 of every warp), profiling builds only (-DLQRO_SIM_PROF,
// tools/simprof.py): [0] findU [1] Jacobian (f) [2] exp: powers + Pade polynomials [3] exp: solve [4] exp: squarings
// [5] noise covariance + Simpson step [6] Jacobi 16x16 + sample [7] propagate tail [8] kalmanFilter1 tail
// [9] observation sample [10] kalmanFilter2 [11] findVGoal [12] load / store [15] agents
#if defined(__CUDACC__)
__device__ long long g_sim_prof[16];
#endif
#if defined(__CUDACC__) && defined(LQRO_SIM_PROF)
#define SIM_TIC() long long sim_t0__ = clock64()
#define SIM_RETIC() sim_t0__ = clock64()
#define SIM_TOC(n) do { const long long t1__ = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)&g_sim_prof[n], (unsigned long long)(t1__ - sim_t0__)); sim_t0__ = t1__; } while (0)
#else
#define SIM_TIC() do { } while (0)
#define SIM_RETIC() do { } while (0)
#define SIM_TOC(n) do { } while (0)
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
  // larger value wins, equal values go to the smaller index: the reference's scans use strict > in increasing
  // index order, i.e. the first maximum.  v >= 0 (an absolute value), so its bit pattern orders like an integer:
  // three REDUX operations (high word, low word among the leaders, index among the ties).
  SIMM static void argmax(double& v, int& idx) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    idx = (int)__reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)idx : 0xffffffffu);
    v = __hiloint2double((int)mh, (int)ml);
  }
};
// hypot(x, y) as MAT:715,717 call it (|x| unbounded, y = 1): one fused x*x + y*y and an IEEE square root, i.e. within
// one ulp of the CRT's result.  Only the Jacobi rotation angles depend on it (no decision does).
SIMF double sim_hypot(double x, double y) { return sqrt(fma(x, x, y * y)); }
#else
typedef SimSeq SimWarp;
SIMF double sim_hypot(double x, double y) { return ::hypot(x, y); }
#endif

#define SIM_FOR(P, e, n) for (int e = P::lane(); e < (n); e += P::stride())

// ---- small dense algebra in the reference's operation order ---------------------------------------
#if defined(__CUDACC__)
// 16x16x16 for one warp: lane = (h, j) owns column j of C for every second row, k outermost: one B operand and four
// 16-byte A loads feed eight independent left-to-right sums (each output still adds its products in k order).
template <bool TB>
SIMF void sim_mm16_warp(double* C, const double* A, const double* B) {
  const int lane = (int)(threadIdx.x & 31), h = lane >> 4, j = lane & 15;
  double acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.0;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const double b0 = TB ? B[j * 16 + k] : B[k * 16 + j];
    const double b1 = TB ? B[j * 16 + k + 1] : B[(k + 1) * 16 + j];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const double2 a = *reinterpret_cast<const double2*>(A + (h + 2 * t) * 16 + k);
      acc[t] += a.x * b0;
      acc[t] += a.y * b1;
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) C[(h + 2 * t) * 16 + j] = acc[t];
  __syncwarp();
}
#endif

// C (n x p) = A (n x m) * B (m x p); TB: B is given transposed (stored p x m).  MAT:219-231.
template <class P, bool TB>
SIMF void sim_mm(double* C, const double* A, const double* B, int n, int m, int p) {
#if defined(__CUDACC__)
  if (P::stride() == 32 && n == 16 && m == 16 && p == 16) { sim_mm16_warp<TB>(C, A, B); return; }  // resolved at compile time after inlining
#endif
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

#if defined(__CUDACC__)
// The same elimination for one warp: lane = (h, c) = (lane >> 4, lane & 15) works on column position c of m and column
// c of inv, for every second row.  No index divisions; a step first gathers its row offsets, factors and operands
// (independent loads), then computes and stores.
template <int N>
SIMF void sim_solve_warp(double* m, double* inv, int* /*rp*/, int* /*cp*/, double* fac) {
  // The row / column permutations live in registers, one nibble per position (every lane keeps the same copy), so a
  // pivot exchange costs no shared-memory round trip; the pivot search of step k+1 is folded into the update of step k
  // (its candidates are exactly the elements that update has just produced).
  const int lane = (int)(threadIdx.x & 31), h = lane >> 4, c = lane & 15;
  const bool act = c < N;
  unsigned long long rpk = 0xfedcba9876543210ull, cpk = 0xfedcba9876543210ull;
#define SIM_NIB(pk, i) ((int)((pk) >> (4 * (i))) & 15)
  double best = 0.0; int key = 0x7fffffff;  // key = (row position << 4) | column position: row-major order of the scan
  if (act) {
    for (int i0 = h; i0 < N; i0 += 8) {
      double a[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { const int i = i0 + 2 * t; a[t] = i < N ? fabs(m[i * N + c]) : 0.0; }
#pragma unroll
      for (int t = 0; t < 4; ++t) { if (a[t] > best) { best = a[t]; key = ((i0 + 2 * t) << 4) | c; } }
    }
  }
  for (int k = 0; k < N; ++k) {
    SimWarp::argmax(best, key);
    if (!(best > 0.0)) key = (k << 4) | k;
    {
      const int mr = key >> 4, mc = key & 15;
      const unsigned long long rk_ = SIM_NIB(rpk, k), rm_ = SIM_NIB(rpk, mr), ck_ = SIM_NIB(cpk, k), cm_ = SIM_NIB(cpk, mc);
      rpk = (rpk & ~(0xfull << (4 * k)) & ~(0xfull << (4 * mr))) | (rm_ << (4 * k)) | (mr != k ? rk_ << (4 * mr) : 0ull);
      cpk = (cpk & ~(0xfull << (4 * k)) & ~(0xfull << (4 * mc))) | (cm_ << (4 * k)) | (mc != k ? ck_ << (4 * mc) : 0ull);
    }
    const int rk = SIM_NIB(rpk, k) * N, ck = SIM_NIB(cpk, k), mcol = act ? SIM_NIB(cpk, c) : 0;
    if (lane > k && lane < N) fac[lane] = m[SIM_NIB(rpk, lane) * N + ck] / m[rk + ck];
    const double pm = m[rk + mcol], pi = act ? inv[rk + c] : 0.0;
    __syncwarp();
    best = 0.0; key = 0x7fffffff;
    if (act) {
      for (int i0 = k + 1 + h; i0 < N; i0 += 8) {  // rows in chunks of four: gather, compute, store
        int rr[4]; double ff[4], mv[4], iv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int i = i0 + 2 * t; const bool ok = i < N; rr[t] = ok ? SIM_NIB(rpk, i) * N : -1; ff[t] = ok ? fac[i] : 0.0; }
#pragma unroll
        for (int t = 0; t < 4; ++t) { if (rr[t] >= 0) { mv[t] = m[rr[t] + mcol]; iv[t] = inv[rr[t] + c]; } }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (rr[t] >= 0) {
            if (c > k) {
              const double nv = mv[t] - ff[t] * pm;
              m[rr[t] + mcol] = nv;
              const double av = fabs(nv);
              if (av > best) { best = av; key = ((i0 + 2 * t) << 4) | c; }  // candidate of the next step's pivot scan
            }
            inv[rr[t] + c] = iv[t] - ff[t] * pi;
          }
        }
      }
    }
    __syncwarp();
  }
  for (int k = N - 1; k >= 0; --k) {
    const int rk = SIM_NIB(rpk, k) * N, ck = SIM_NIB(cpk, k);
    const double pv = act ? inv[rk + c] / m[rk + ck] : 0.0;  // both half-warps form the same quotient
    __syncwarp();
    if (act) {
      if (h == 0) inv[rk + c] = pv;
      for (int i0 = h; i0 < k; i0 += 8) {
        int rr[4]; double mf[4], iv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int i = i0 + 2 * t; rr[t] = i < k ? SIM_NIB(rpk, i) * N : -1; }
#pragma unroll
        for (int t = 0; t < 4; ++t) { if (rr[t] >= 0) { mf[t] = m[rr[t] + ck]; iv[t] = inv[rr[t] + c]; } }
#pragma unroll
        for (int t = 0; t < 4; ++t) { if (rr[t] >= 0) inv[rr[t] + c] = iv[t] - mf[t] * pv; }
      }
    }
    __syncwarp();
  }
  if (act) {
    for (int i = h; i < N; i += 2) m[SIM_NIB(cpk, i) * N + c] = inv[SIM_NIB(rpk, i) * N + c];
  }
  __syncwarp();
#undef SIM_NIB
}
template <> SIMF void sim_solve<SimWarp, 16>(double* m, double* inv, int* rp, int* cp, double* fac) { sim_solve_warp<16>(m, inv, rp, cp, fac); }
template <> SIMF void sim_solve<SimWarp, 6>(double* m, double* inv, int* rp, int* cp, double* fac) { sim_solve_warp<6>(m, inv, rp, cp, fac); }
#endif

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

// 1-norm of q (column sums top to bottom, MAT:275-288) and the scaling exponent of exp (MAT:776)
template <class P, int N>
SIMF double sim_norm1(const double* q, double* small) {
  SIM_FOR(P, j, N) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += fabs(q[i * N + j]);
    small[j] = s;
  }
  P::sync();
  double norm1 = 0.0;
  for (int j = 0; j < N; ++j) if (small[j] > norm1) norm1 = small[j];
  P::sync();  // small[] is reused by the solve
  return norm1;
}
SIMF int sim_exp_scaling(double norm1) {
  const double sd = ceil(log(norm1 / SIM_NORMLIM) * SIM_LOG2E);
  return (int)(sd > 0.0 ? sd : 0.0);
}
// The part of exp(q) before the squarings (MAT:779-786): q is scaled IN PLACE by 2^-s (destroyed), R7 = (V - U) % (V + U)
// is left in ws[N*N .. 2N*N) ("A4"); ws[0 .. N*N) ("A2") is free afterwards.  ws: 4*N*N doubles, perm: 2N ints, small: N.
template <class P, int N>
SIMF void sim_exp_pade(double* q, int s, double* ws, int* perm, double* small) {
  double* A = q; double* A2 = ws; double* A4 = ws + N * N; double* A6 = ws + 2 * N * N; double* T = ws + 3 * N * N;
  SIM_TIC();
  const double rscale = ldexp(1.0, -s);  // A /= 2^s (MAT:779): multiplying by 2^-s gives the same bits (both exact up to underflow)
  SIM_FOR(P, e, N * N) A[e] = A[e] * rscale;
  P::sync();
  sim_mm<P, false>(A2, A, A, N, N, N);
  sim_mm<P, false>(A4, A2, A2, N, N, N);
  sim_mm<P, false>(A6, A2, A4, N, N, N);
  SIM_FOR(P, e, N * N) {  // the two Pade polynomials: U's factor into T, V over A6
    const double id = (e / N == e % N) ? 1.0 : 0.0;
    const double a6 = A6[e], a4 = A4[e], a2 = A2[e];
    T[e] = a6 * SIM_JB7 + a4 * SIM_JB5 + a2 * SIM_JB3 + id * SIM_JB1;
    A6[e] = a6 * SIM_JB6 + a4 * SIM_JB4 + a2 * SIM_JB2 + id * SIM_JB0;
  }
  P::sync();
  sim_mm<P, false>(A2, A, T, N, N, N);  // U (A2 is dead)
  SIM_FOR(P, e, N * N) { const double v = A6[e], u = A2[e]; A4[e] = v - u; T[e] = v + u; }
  P::sync();
  if (N == 16) SIM_TOC(2);
  sim_solve<P, N>(A4, T, perm, perm + N, small);  // R7 = (V - U) % (V + U), result in A4
  if (N == 16) SIM_TOC(3);
}
// out = exp(q) by scaling, the (7,7) Pade approximant and squaring (MAT:773-790).  q is scaled IN PLACE by 2^-s (s is
// returned); ws: 4*N*N doubles, perm: 2N ints, small: N doubles.  out may not alias q or the first two buffers of ws (the
// last two, A6 and T, are dead when out is written).
template <class P, int N>
SIMF int sim_exp(double* out, double* q, double* ws, int* perm, double* small) {
  const int s = sim_exp_scaling(sim_norm1<P, N>(q, small));
  sim_exp_pade<P, N>(q, s, ws, perm, small);
  SIM_TIC();
  double* cur = ws + N * N; double* nxt = ws;
  for (int i = 0; i < s; ++i) {
    sim_mm<P, false>(nxt, cur, cur, N, N, N);
    double* t = cur; cur = nxt; nxt = t;
  }
  SIM_FOR(P, e, N * N) out[e] = cur[e];
  P::sync();
  if (N == 16) SIM_TOC(4);
  return s;
}
// full = exp(q) and half = exp(q/2) together -- propagate and kalmanFilter1 ask for exp(dt*F) and exp((dt*0.5)*F)
// (MAIN:446-447, 461-462), and (dt*0.5)*F is dt*F halved bit for bit (halving is exact).  exp scales its argument by 2^-s
// with s = max(0, ceil(log2(|q|_1 / 0.95))) (MAT:776-779); |q/2|_1 = |q|_1 / 2 exactly, so s(q/2) = s(q) - 1 whenever
// s(q) >= 1 (checked with the reference's own formula on both norms): both exponentials then start from THE SAME scaled
// matrix, hence the same A^2, A^4, A^6, U, V and R7, and differ only in the number of squarings -- half = R7^(2^(s-1)),
// full = half*half, each the very product the reference forms.  Exact work avoidance: one Pade approximant and one
// full-pivot solve instead of two.  Returns false, with q untouched, when the exponents do not line up (s(q) = 0).
// full / half may be the last two buffers of ws (dead by then), not q or the first two.
template <class P, int N>
SIMF bool sim_exp_with_half(double* full, double* half, double* q, double* ws, int* perm, double* small) {
  const double norm1 = sim_norm1<P, N>(q, small);
  const int s = sim_exp_scaling(norm1), sh = sim_exp_scaling(norm1 * 0.5);
  if (s < 1 || sh != s - 1) return false;
  sim_exp_pade<P, N>(q, s, ws, perm, small);
  SIM_TIC();
  double* cur = ws + N * N; double* nxt = ws;
  for (int i = 0; i < sh; ++i) {
    sim_mm<P, false>(nxt, cur, cur, N, N, N);
    double* t = cur; cur = nxt; nxt = t;
  }
  SIM_FOR(P, e, N * N) half[e] = cur[e];
  sim_mm<P, false>(full, cur, cur, N, N, N);
  if (N == 16) SIM_TOC(4);
  return true;
}

#ifdef SIM_COUNT_JACOBI
static long sim_jacobi_counts[4];  // host-side instrumentation: iterations / rotations of the 16x16 and 6x6 problems
#endif
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
#ifdef SIM_COUNT_JACOBI
    sim_jacobi_counts[N == 16 ? 0 : 2] += 1; if (maximum > DBL_EPSILON) sim_jacobi_counts[N == 16 ? 1 : 3] += 1;
#endif
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
#pragma unroll
  for (int i = 0; i < 3; ++i) { double t = 0.0; t += Mx[i * 3] * v[0]; t += Mx[i * 3 + 1] * v[1]; t += Mx[i * 3 + 2] * v[2]; y[i] = t; }
}
SIMF void sim_mtv3(double* y, const double* Mx, const double* v) {  // (~M) * v
#pragma unroll
  for (int i = 0; i < 3; ++i) { double t = 0.0; t += Mx[i] * v[0]; t += Mx[3 + i] * v[1]; t += Mx[6 + i] * v[2]; y[i] = t; }
}
SIMF double sim_hypot3(const double* v) {  // stdafx.h:86-97
  const double X = fabs(v[0]), Y = fabs(v[1]), Z = fabs(v[2]);
  if (X > Y && X > Z) return X * sqrt(1.0 + (Y / X) * (Y / X) + (Z / X) * (Z / X));
  else if (Y > Z) return Y * sqrt(1.0 + (X / Y) * (X / Y) + (Z / Y) * (Z / Y));
  else return Z * sqrt(1.0 + (X / Z) * (X / Z) + (Y / Z) * (Y / Z));
}
// 3x3 product, fully unrolled (every index a compile-time constant: the operands stay in registers)
SIMF void sim_mm3(double* C, const double* A, const double* B) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) t += A[i * 3 + k] * B[k * 3 + j];
      C[i * 3 + j] = t;
    }
  }
}
// P X = Q for 3x3 operands in registers: the full-pivot elimination of MAT:370-440 with the pivot brought to (k, k) by
// physically exchanging rows / columns (the reference permutes indices instead; the arithmetic on each logical element
// and the row-major order of the pivot scan over POSITIONS are the same).  m: in P, destroyed; inv: in Q, destroyed;
// X: out.
SIMF void sim_solve3(double* X, double* m, double* inv) {
  int cp0 = 0, cp1 = 1, cp2 = 2;  // logical column at each column position
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double best = 0.0; int mr = k, mc = k;
#pragma unroll
    for (int i = k; i < 3; ++i) {
#pragma unroll
      for (int j = k; j < 3; ++j) {
        const double a = fabs(m[i * 3 + j]);
        if (a > best) { best = a; mr = i; mc = j; }
      }
    }
#pragma unroll
    for (int r = k + 1; r < 3; ++r) {
      if (mr == r) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double t = m[k * 3 + j]; m[k * 3 + j] = m[r * 3 + j]; m[r * 3 + j] = t;
          t = inv[k * 3 + j]; inv[k * 3 + j] = inv[r * 3 + j]; inv[r * 3 + j] = t;
        }
      }
    }
#pragma unroll
    for (int c = k + 1; c < 3; ++c) {
      if (mc == c) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double t = m[i * 3 + k]; m[i * 3 + k] = m[i * 3 + c]; m[i * 3 + c] = t; }
        if (k == 0) { const int t = cp0; if (c == 1) { cp0 = cp1; cp1 = t; } else { cp0 = cp2; cp2 = t; } }
        else { const int t = cp1; cp1 = cp2; cp2 = t; }
      }
    }
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      const double factor = m[i * 3 + k] / m[k * 3 + k];
#pragma unroll
      for (int j = k + 1; j < 3; ++j) m[i * 3 + j] -= factor * m[k * 3 + j];
#pragma unroll
      for (int j = 0; j < 3; ++j) inv[i * 3 + j] -= factor * inv[k * 3 + j];
    }
  }
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const double quotient = m[k * 3 + k];
#pragma unroll
    for (int j = 0; j < 3; ++j) inv[k * 3 + j] /= quotient;
#pragma unroll
    for (int i = 0; i < k; ++i) {
      const double factor = m[i * 3 + k];
#pragma unroll
      for (int j = 0; j < 3; ++j) inv[i * 3 + j] -= factor * inv[k * 3 + j];
    }
  }
  // X(cp[i], :) = row i
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int j = 0; j < 3; ++j) X[r * 3 + j] = (cp0 == r) ? inv[j] : ((cp1 == r) ? inv[3 + j] : inv[6 + j]);
  }
}
// exp of a 3x3 matrix (MAT:773-790) with everything in registers
SIMF void sim_exp3(double* out, const double* q) {
  double norm1 = 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) c += fabs(q[i * 3 + j]);
    if (c > norm1) norm1 = c;
  }
  const double sd = ceil(log(norm1 / SIM_NORMLIM) * SIM_LOG2E);
  const int s = (int)(sd > 0.0 ? sd : 0.0);
  const double rscale = ldexp(1.0, -s);  // A /= 2^s: the product with 2^-s has the same bits
  double A[9], A2[9], A4[9], A6[9], T[9], U[9], V[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) A[e] = q[e] * rscale;
  sim_mm3(A2, A, A);
  sim_mm3(A4, A2, A2);
  sim_mm3(A6, A2, A4);
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    const double id = (e % 4 == 0) ? 1.0 : 0.0;
    T[e] = A6[e] * SIM_JB7 + A4[e] * SIM_JB5 + A2[e] * SIM_JB3 + id * SIM_JB1;
    V[e] = A6[e] * SIM_JB6 + A4[e] * SIM_JB4 + A2[e] * SIM_JB2 + id * SIM_JB0;
  }
  sim_mm3(U, A, T);
#pragma unroll
  for (int e = 0; e < 9; ++e) { A2[e] = V[e] - U[e]; A4[e] = V[e] + U[e]; }
  sim_solve3(out, A2, A4);
  for (int i = 0; i < s; ++i) {
    sim_mm3(T, out, out);
#pragma unroll
    for (int e = 0; e < 9; ++e) out[e] = T[e];
  }
}
// R <- R * exp(skew(r))   (the "reset rotation-error" step, MAIN:453, 472, 485)
SIMF void sim_fold_rotation(double* R, const double* r) {
  double S[9], Ex[9], Rn[9];
  sim_skew(S, r);
  sim_exp3(Ex, S);
  sim_mm3(Rn, R, Ex);
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
    sim_mm3(RE, R, Ex);
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
    sim_mm3(SJ, S, md.inertia);
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
  sim_mm3(RLocal, Ex, R0);
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

// ---- per-agent (per-warp) workspaces of the three phases --------------------------------------------------
// A: one linearised flow (propagate's about the true state, or kalmanFilter1's about the estimate): FIVE 16x16 buffers whose
// roles rotate through the phase (see sim_flow), 11 KB per warp
struct alignas(16) LqroSimWsA {
  double buf[5 * 256];
  double xs[16], Rs[9], u[4], xdot[16], v1[16], small[16];
  int perm[32], m_diag, pad_;  // m_diag: the motion-noise covariance M is diagonal (set by the caller; see sim_flow)
  SIMM double* A2e() { return buf + 3 * 256; }  // after sim_flow: exp(dt/2 F)
  SIMM double* Ae() { return buf + 4 * 256; }   //                 exp(dt F)
  SIMM double* MM() { return buf + 1 * 256; }   //                 the noise covariance
};
// B: the process-noise sample (Jacobi eigen-decomposition of propagate's MM)
struct alignas(16) LqroSimWsB {
  double D[256], V[256], nrm[16], samp[16], zero[16];
};
// C: kalmanFilter1's and propagate's state updates, the observation, kalmanFilter2, the next goal velocity
struct alignas(16) LqroSimWsC {
  double P[256], W[4 * 256];
  double x[16], xT[16], R[9], RT[9], z[6], hz[6], nrm[6], v1[16], v2[16], small[16];
  int perm[32];
};
// 6-dimensional operands of the observation / kalmanFilter2 overlay the part of W the 16x16 products there leave free
struct LqroSimObs {
  double *H, *PHt, *HP, *S6, *S6b, *K, *V6;
  SIMM explicit LqroSimObs(double* W) : H(W + 512), PHt(W + 608), HP(W + 704), S6(W + 800), S6b(W + 836), K(W + 872), V6(W + 968) {}
};

#if defined(__CUDACC__)
#define SIM_FORCE_GENERAL_EXP 0
#else
static int sim_force_general_exp = 0;  // CPU build only: lets the tests run sim_flow's general path on inputs that never take it
#define SIM_FORCE_GENERAL_EXP sim_force_general_exp
#endif
// Common half of propagate and kalmanFilter1 (MAIN:443-450, 458-468): F = Jacobian_fx, xDot = f, A = exp(dt F),
// A2 = exp(dt/2 F), MM = (dt/6)(M + 4 A2 M ~A2 + A M ~A), v1 = (dt/6)(xDot + 4 (A2 xDot) + A xDot).
// xs, Rs: the state / rotation linearised about (shared memory of the warp).  `park`: 256 doubles outside the workspace
// (the task's output slot in HBM), used only by the general path of the exponentials.
// Buffer roles (B0..B4 = w.buf):  Jacobian: f staging in B1..B3, F -> B0;  q = dt F in place in B0;  exponentials: A = B0,
// A2 A4 A6 T = B1..B4, exp(dt/2 F) -> B3, exp(dt F) -> B4;  noise covariance: temporaries B0, B1, B2, MM -> B1.
template <class P>
SIMN void sim_flow(LqroSimWsA& w, const double* xs, const double* Rs, double* park, const LqroSimModel& md) {
  SIM_TIC();
  double* const B0 = w.buf; double* const B1 = w.buf + 256; double* const B2 = w.buf + 512;
  double* const A2e = w.A2e(); double* const Ae = w.Ae();
  // Central-difference Jacobian (MAIN:392-401): the 32 evaluations f(x +- jStep e_c) run one per lane, results staged.
  // sim_f never reads the position x[0..2] (MAIN:341 extracts p and does not use it), so f(x + jStep e_0) is f(x) bit for
  // bit: the unperturbed xDot (MAIN:444) is column 0's evaluation.
  double* fb = B1;  // [32][17]: padded rows, conflict-free for lane-per-row stores
  SIM_FOR(P, t, 32) {
    const int c = t & 15, side = t >> 4;
    double xr[16], R[9], u[4], ft[16];
    for (int i = 0; i < 9; ++i) R[i] = Rs[i];
    for (int i = 0; i < 4; ++i) u[i] = w.u[i];
    for (int i = 0; i < 16; ++i) xr[i] = xs[i];
    for (int i = 0; i < 16; ++i) if (i == c) { if (side == 0) xr[i] += md.jstep; else xr[i] -= md.jstep; }
    sim_f(ft, xr, R, u, md);
    for (int i = 0; i < 16; ++i) fb[t * 17 + i] = ft[i];
  }
  P::sync();
  SIM_FOR(P, e, 256) {  // q = dt*F directly (MAIN:446: exp(dt*F)); F itself is not needed again
    const int i = e >> 4, c = e & 15;
    B0[e] = ((fb[c * 17 + i] - fb[(16 + c) * 17 + i]) / (2 * md.jstep)) * md.dt;
  }
  SIM_FOR(P, i, 16) w.xdot[i] = fb[i];
  P::sync();
  SIM_TOC(1);
  if (SIM_FORCE_GENERAL_EXP || !sim_exp_with_half<P, 16>(Ae, A2e, B0, B1, w.perm, w.small)) {
    // general path: two separate exponentials.  exp(q) is parked outside the workspace while exp(q/2) is formed; q/2 is
    // recovered from the scaled copy exp left in B0 (q 2^-s, exact): q/2 = (q 2^-s) 2^(s-1)
    const int s1 = sim_exp<P, 16>(park, B0, B1, w.perm, w.small);
    const double back = ldexp(1.0, s1 - 1);
    SIM_FOR(P, e, 256) B0[e] = B0[e] * back;
    P::sync();
    sim_exp<P, 16>(A2e, B0, B1, w.perm, w.small);
    SIM_FOR(P, e, 256) Ae[e] = park[e];
    P::sync();
  }
  // MM = (dt/6) * (M + 4*A2*M*~A2 + A*M*~A)
  SIM_RETIC();
  SIM_FOR(P, e, 256) B0[e] = A2e[e] * 4;
  P::sync();
  // X*M for a DIAGONAL M (the reference's M = 1e-9 I, MAIN:1139): the 16-term sum 0 + X(i,0) M(0,j) + ... has one non-zero
  // term, so it equals 0.0 + X(i,j) M(j,j) bit for bit (the leading +0 reproduces the sum's +0 when the product is -0)
  if (w.m_diag) { SIM_FOR(P, e, 256) B1[e] = 0.0 + B0[e] * md.M[(e & 15) * 17]; P::sync(); }
  else sim_mm<P, false>(B1, B0, md.M, 16, 16, 16);
  sim_mm<P, true>(B2, B1, A2e, 16, 16, 16);   // 4*A2*M*~A2
  if (w.m_diag) { SIM_FOR(P, e, 256) B1[e] = 0.0 + Ae[e] * md.M[(e & 15) * 17]; P::sync(); }
  else sim_mm<P, false>(B1, Ae, md.M, 16, 16, 16);
  sim_mm<P, true>(B0, B1, Ae, 16, 16, 16);    // A*M*~A
  SIM_FOR(P, e, 256) B1[e] = (md.M[e] + B2[e] + B0[e]) * (md.dt / 6);
  // v1 = (dt/6)*(xDot + 4*(A2*xDot) + A*xDot)
  SIM_FOR(P, i, 16) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 16; ++k) a += A2e[i * 16 + k] * w.xdot[k];
    for (int k = 0; k < 16; ++k) b += Ae[i * 16 + k] * w.xdot[k];
    w.v1[i] = (w.xdot[i] + a * 4 + b) * (md.dt / 6);
  }
  P::sync();
  SIM_TOC(5);
}

// One agent, MAIN:1409-1416, in phases (four kernels on the device, so that each runs at the occupancy its own
// shared-memory footprint allows; the host build calls them back to back):
//   control  u = findU()                                                        thread per agent
//   flow     the linearised flow about the TRUE state (propagate, MAIN:443-450) and about the ESTIMATE
//            (kalmanFilter1, MAIN:458-468): two independent tasks per agent       warp per task
//   noise    sampleGaussian(0, MM) of propagate (MAIN:450)                       warp per agent
//   filter   the state / covariance updates of kalmanFilter1 and propagate, the observation, kalmanFilter2, findVGoal
// kalmanFilter1 reads nothing propagate writes (and vice versa), so their order does not change any value.
//
// Flow phase.  In: w.xs, w.Rs (the state / rotation linearised about), w.u.  Out: MM[256], v1[16] (noise covariance and
// Simpson increment) and, if Ae_out != 0, A = exp(dt F)[256].
template <class P>
SIMF void sim_phase_flow(LqroSimWsA& w, double* MM_out, double* v1_out, double* Ae_out, const LqroSimModel& md) {
  sim_flow<P>(w, w.xs, w.Rs, MM_out, md);
  SIM_FOR(P, e, 256) MM_out[e] = w.MM()[e];
  SIM_FOR(P, i, 16) v1_out[i] = w.v1[i];
  if (Ae_out) { SIM_FOR(P, e, 256) Ae_out[e] = w.Ae()[e]; }
  P::sync();
}

// Quadrotor::setupQuadrotors' noise part (MAIN:126-129) and the first findVGoal (MAIN:1355): xTrue = sampleGaussian(x, P);
// RotTrue = Rot*exp([xTrue_r]); xTrue_r = 0; vGoal = findVGoal().  In: w.D = P, w.nrm = 16 normals, w.zero = x (the mean).
template <class P>
SIMF void sim_phase_init(LqroSimWsB& w, const double* x, const double* Rot, const double* pGoal, double* xT, double* RT, double* vG,
                         const LqroSimModel& md) {
  sim_sample_gaussian<P, 16>(w.samp, w.zero, w.D, w.V, w.nrm);
  double r[3], R[9];
  for (int i = 0; i < 3; ++i) r[i] = w.samp[6 + i];
  for (int i = 0; i < 9; ++i) R[i] = Rot[i];
  sim_fold_rotation(R, r);
  for (int i = 0; i < 16; ++i) xT[i] = (i >= 6 && i < 9) ? 0.0 : w.samp[i];
  for (int i = 0; i < 9; ++i) RT[i] = R[i];
  sim_find_vgoal(vG, x, Rot, pGoal, md);
}

// linearizeDiscretize (MAIN:425-440) about (w.xs, w.Rs, w.u): A = exp(dt F), B = Int*G, c = Int*xDot with
// Int = (dt/6)(I + 4 exp(dt/2 F) + A) and G = Jacobian_fu (central differences in the four inputs, MAIN:403-412).
// The flow phase already forms F, xDot, A and exp(dt/2 F); with lqro_solve_gains this is controlMatrices (MAIN:490-544).
template <class P>
SIMF void sim_phase_linearize(LqroSimWsA& w, double* A_out, double* B_out, double* c_out, const LqroSimModel& md) {
  sim_flow<P>(w, w.xs, w.Rs, A_out, md);
  const double* A2e = w.A2e(); const double* Ae = w.Ae();
  double* fb = w.buf; double* G = w.buf + 160; double* Int = w.buf + 512;  // fb [8][17], G 16x4 (B0), Int 16x16 (B2)
  SIM_FOR(P, t, 8) {
    const int c = t & 3, side = t >> 2;
    double xr[16], R[9], u[4], ft[16];
    for (int i = 0; i < 16; ++i) xr[i] = w.xs[i];
    for (int i = 0; i < 9; ++i) R[i] = w.Rs[i];
    for (int i = 0; i < 4; ++i) u[i] = w.u[i];
    for (int i = 0; i < 4; ++i) if (i == c) { if (side == 0) u[i] += md.jstep; else u[i] -= md.jstep; }
    sim_f(ft, xr, R, u, md);
    for (int i = 0; i < 16; ++i) fb[t * 17 + i] = ft[i];
  }
  P::sync();
  SIM_FOR(P, e, 64) { const int i = e >> 2, c = e & 3; G[e] = (fb[c * 17 + i] - fb[(4 + c) * 17 + i]) / (2 * md.jstep); }
  SIM_FOR(P, e, 256) { const double id = ((e >> 4) == (e & 15)) ? 1.0 : 0.0; Int[e] = (id + A2e[e] * 4.0 + Ae[e]) * (md.dt / 6.0); }
  P::sync();
  sim_mm<P, false>(B_out, Int, G, 16, 16, 4);
  SIM_FOR(P, i, 16) { double t = 0.0; for (int k = 0; k < 16; ++k) t += Int[i * 16 + k] * w.xdot[k]; c_out[i] = t; }
  SIM_FOR(P, e, 256) A_out[e] = Ae[e];
  P::sync();
}

// Phase B.  In: w.D = propagate's MM, w.nrm = the 16 process-noise normals.  Out: w.samp = sampleGaussian(0, MM)
// (MAIN:450, SIM2:19-31).
template <class P>
SIMF void sim_phase_b(LqroSimWsB& w) {
  SIM_TIC();
  SIM_FOR(P, i, 16) w.zero[i] = 0.0;
  P::sync();
  sim_sample_gaussian<P, 16>(w.samp, w.zero, w.D, w.V, w.nrm);
  SIM_TOC(6);
}

// Filter phase.  In: w.x, w.R, w.P (estimate), w.xT, w.RT (true state), the flow results of both tasks (Ae_k, MM_k, v1k:
// kalmanFilter1's; v1p: propagate's), samp (noise phase), w.nrm = the 6 observation normals, pGoal.  Finishes
// kalmanFilter1 (MAIN:465-473) and propagate (MAIN:450-454), draws the observation (MAIN:1413), runs kalmanFilter2
// (MAIN:476-487) and findVGoal (MAIN:1416).
template <class P>
SIMF void sim_phase_filter(LqroSimWsC& w, const double* Ae_k, const double* MM_k, const double* v1k, const double* v1p,
                           const double* samp, const double* pGoal, double* vG, const LqroSimModel& md) {
  SIM_TIC();
  {  // kalmanFilter1: P = A*P*~A + MM; x += increment; fold the rotation error
    double* T1 = w.W; double* T2 = w.W + 256; double* As = w.W + 512;
    SIM_FOR(P, e, 256) As[e] = Ae_k[e];  // stage A (HBM scratch of the flow kernel) once: both products read it 16 times over
    P::sync();
    sim_mm<P, false>(T1, As, w.P, 16, 16, 16);
    sim_mm<P, true>(T2, T1, As, 16, 16, 16);
    SIM_FOR(P, e, 256) w.P[e] = T2[e] + MM_k[e];
    double r[3], R[9];
    for (int i = 0; i < 3; ++i) r[i] = w.x[6 + i] + v1k[6 + i];
    for (int i = 0; i < 9; ++i) R[i] = w.R[i];
    sim_fold_rotation(R, r);
    P::sync();
    SIM_FOR(P, i, 16) w.x[i] = (i >= 6 && i < 9) ? 0.0 : w.x[i] + v1k[i];
    SIM_FOR(P, i, 9) w.R[i] = R[i];
    P::sync();
  }
  SIM_TOC(8);
  {
    double r[3], R[9];
    for (int i = 0; i < 3; ++i) r[i] = w.xT[6 + i] + v1p[6 + i] + samp[6 + i];
    for (int i = 0; i < 9; ++i) R[i] = w.RT[i];
    sim_fold_rotation(R, r);
    P::sync();
    SIM_FOR(P, i, 16) w.xT[i] = (i >= 6 && i < 9) ? 0.0 : w.xT[i] + v1p[i] + samp[i];
    SIM_FOR(P, i, 9) w.RT[i] = R[i];
    P::sync();
  }
  SIM_TOC(7);
  const LqroSimObs o(w.W);
  // Observation z = sampleGaussian(h(xTrue, RotTrue), N)  (MAIN:1413; h MAIN:369-390: rate gyros, then position)
  {
    SIM_FOR(P, i, 6) w.hz[i] = i < 3 ? w.xT[9 + i] : w.xT[i - 3];
    SIM_FOR(P, e, 36) o.S6[e] = md.N[e];
    P::sync();
    sim_sample_gaussian<P, 6>(w.z, w.hz, o.S6, o.V6, w.nrm);
  }
  SIM_TOC(9);
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
  SIM_TOC(10);
  {  // vGoal = findVGoal()  (MAIN:1416)
    double x[16], R[9];
    for (int i = 0; i < 16; ++i) x[i] = w.x[i];
    for (int i = 0; i < 9; ++i) R[i] = w.R[i];
    sim_find_vgoal(vG, x, R, pGoal, md);
  }
  SIM_TOC(11);
}

#if defined(__CUDACC__)
// Scratch in HBM between the kernels (per agent): u[4]; per flow task MM[256] + v1[16]; kalmanFilter1's A[256]; samp[16].
struct LqroSimScratch {
  double *u, *MM, *v1, *Ae, *samp;  // u [n][4], MM [2n][256], v1 [2n][16], Ae [n][256], samp [n][16]; task = 2*agent + {0: true state, 1: estimate}
};

__global__ void __launch_bounds__(128) k_sim_control(int n, const LqroSimModel* __restrict__ model, const double* __restrict__ x,
                                                     const double* __restrict__ Rot, const double* __restrict__ newV,
                                                     double* __restrict__ u, double* __restrict__ uOut) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  double xs[16], R[9], nv[3], uu[4];
  for (int i = 0; i < 16; ++i) xs[i] = x[(size_t)a * 16 + i];
  for (int i = 0; i < 9; ++i) R[i] = Rot[(size_t)a * 9 + i];
  for (int i = 0; i < 3; ++i) nv[i] = newV[(size_t)a * 3 + i];
  sim_find_u(uu, xs, R, nv, *model);
  for (int i = 0; i < 4; ++i) { u[(size_t)a * 4 + i] = uu[i]; if (uOut) uOut[(size_t)a * 4 + i] = uu[i]; }
}

// initial true states and goal velocities (warp per agent; runs once)
__global__ void __launch_bounds__(128) k_sim_init(int n, const LqroSimModel* __restrict__ model, const double* __restrict__ x,
                                                  const double* __restrict__ Rot, const double* __restrict__ Pcov,
                                                  const double* __restrict__ pGoal, const double* __restrict__ normals16,
                                                  double* __restrict__ xTrue, double* __restrict__ RotTrue, double* __restrict__ vGoal) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  const LqroSimModel& md = *model;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  LqroSimWsB& w = *reinterpret_cast<LqroSimWsB*>(sim_smem + (size_t)warp * sizeof(LqroSimWsB));
  for (int a = blockIdx.x * wpb + warp; a < n; a += gridDim.x * wpb) {
    for (int e = lane; e < 256; e += 32) w.D[e] = Pcov[(size_t)a * 256 + e];
    if (lane < 16) { w.nrm[lane] = normals16[(size_t)a * 16 + lane]; w.zero[lane] = x[(size_t)a * 16 + lane]; }
    __syncwarp();
    double xs[16], R[9], pg[3], xT[16], RT[9], vG[3];
    for (int i = 0; i < 16; ++i) xs[i] = x[(size_t)a * 16 + i];
    for (int i = 0; i < 9; ++i) R[i] = Rot[(size_t)a * 9 + i];
    for (int i = 0; i < 3; ++i) pg[i] = pGoal[(size_t)a * 3 + i];
    sim_phase_init<SimWarp>(w, xs, R, pg, xT, RT, vG, md);
    if (lane == 0) {
      for (int i = 0; i < 16; ++i) xTrue[(size_t)a * 16 + i] = xT[i];
      for (int i = 0; i < 9; ++i) RotTrue[(size_t)a * 9 + i] = RT[i];
      for (int i = 0; i < 3; ++i) vGoal[(size_t)a * 3 + i] = vG[i];
    }
    __syncwarp();
  }
}

// launch geometry of the flow kernels: warps per CTA x CTAs per SM (11 KB of shared memory per warp; the register cap that
// follows is 65536 / (32 x warps x CTAs)): measured at 65,536 agents 5 x 3 (128 registers) 6.0 ms, 5 x 4 (96) 6.55 ms
#ifndef SIM_FLOW_WARPS
#define SIM_FLOW_WARPS 4
#define SIM_FLOW_CTAS 4
#endif
// linearizeDiscretize for a batch of operating points (warp per problem)
__global__ void __launch_bounds__(SIM_FLOW_WARPS * 32, SIM_FLOW_CTAS) k_sim_linearize(int n, const LqroSimModel* __restrict__ model, const double* __restrict__ x,
                                                          const double* __restrict__ Rot, const double* __restrict__ u,
                                                          double* __restrict__ A, double* __restrict__ B, double* __restrict__ c, int m_diag) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  const LqroSimModel& md = *model;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  LqroSimWsA& w = *reinterpret_cast<LqroSimWsA*>(sim_smem + (size_t)warp * sizeof(LqroSimWsA));
  for (int a = blockIdx.x * wpb + warp; a < n; a += gridDim.x * wpb) {
    if (lane < 16) w.xs[lane] = x[(size_t)a * 16 + lane];
    if (lane < 9) w.Rs[lane] = Rot[(size_t)a * 9 + lane];
    if (lane < 4) w.u[lane] = u[(size_t)a * 4 + lane];
    if (lane == 0) w.m_diag = m_diag;
    __syncwarp();
    sim_phase_linearize<SimWarp>(w, A + (size_t)a * 256, B + (size_t)a * 64, c + (size_t)a * 16, md);
  }
}

__global__ void __launch_bounds__(SIM_FLOW_WARPS * 32, SIM_FLOW_CTAS) k_sim_flow(int n, const LqroSimModel* __restrict__ model, const double* __restrict__ x,
                                                     const double* __restrict__ Rot, const double* __restrict__ xTrue,
                                                     const double* __restrict__ RotTrue, LqroSimScratch sc, int m_diag) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  const LqroSimModel& md = *model;  // 3.6 KB of constants and gains, read through L1 (uniform or row-contiguous accesses)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  LqroSimWsA& w = *reinterpret_cast<LqroSimWsA*>(sim_smem + (size_t)warp * sizeof(LqroSimWsA));
  for (int t = blockIdx.x * wpb + warp; t < 2 * n; t += gridDim.x * wpb) {
    const int a = t >> 1, est = t & 1;
    const double* xs = est ? x : xTrue; const double* Rs = est ? Rot : RotTrue;
    if (lane < 16) w.xs[lane] = xs[(size_t)a * 16 + lane];
    if (lane < 9) w.Rs[lane] = Rs[(size_t)a * 9 + lane];
    if (lane < 4) w.u[lane] = sc.u[(size_t)a * 4 + lane];
    if (lane == 0) w.m_diag = m_diag;
    __syncwarp();
    sim_phase_flow<SimWarp>(w, sc.MM + (size_t)t * 256, sc.v1 + (size_t)t * 16, est ? sc.Ae + (size_t)a * 256 : nullptr, md);
  }
}

// The noise phase for FOUR agents per warp (8 lanes each).  The Jacobi iteration (MAT:674-757) is a chain of scalar FP64
// divisions and square roots that every lane of a warp would execute redundantly for ONE agent; with four agents side
// by side the same instructions serve four.  The four groups run the loop in lockstep (a group that has converged idles
// on dummy operands), every group performs exactly the operation sequence of sim_jacobi / sim_sample_gaussian.
// Shared memory per group: D[256] | Vt[256] (eigenvectors TRANSPOSED: a rotation touches two contiguous rows) | nrm[16].
// D is symmetric and the iteration only ever touches its upper triangle (MAT:686-740): packed storage, 136 doubles,
// element (r, c), r <= c, at r(31 - r)/2 + c -- rows of decreasing length, so column accesses spread over the banks
#define SIM_TRI(r, c) ((((r) * (31 - (r))) >> 1) + (c))
#define SIM_NOISE_GROUP_DOUBLES (136 + 256 + 16 + 4)  // +4: group bases fall on different banks
__global__ void __launch_bounds__(128, 4) k_sim_noise(int n, const double* __restrict__ normals, LqroSimScratch sc) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  constexpr int N = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int g = lane >> 3, sl = lane & 7;
  double* D = reinterpret_cast<double*>(sim_smem) + (size_t)(warp * 4 + g) * SIM_NOISE_GROUP_DOUBLES;
  double* Vt = D + 136; double* nrm = Vt + 256;
  for (int a0 = (blockIdx.x * wpb + warp) * 4; a0 < n; a0 += gridDim.x * wpb * 4) {
    const int a = a0 + g;
    bool done = a >= n;
    if (!done) {
      for (int e = sl; e < 256; e += 8) {
        const int r = e >> 4, c = e & 15;
        if (r <= c) D[SIM_TRI(r, c)] = sc.MM[(size_t)(2 * a) * 256 + e];
        Vt[e] = (r == c) ? 1.0 : 0.0;
      }
      for (int e = sl; e < 16; e += 8) nrm[e] = normals[(size_t)a * 22 + e];
    }
    __syncwarp();
    int pivotRow = 0, zeroCount = 0;
    while (!__all_sync(0xffffffffu, done)) {
      double maximum = 0.0; int bi = 0x7fffffff;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int k = sl + 8 * t;
        if (k != pivotRow) {
          const double av = fabs(k < pivotRow ? D[SIM_TRI(k, pivotRow)] : D[SIM_TRI(pivotRow, k)]);
          if (av > maximum) { maximum = av; bi = k; }
        }
      }
#pragma unroll
      for (int o = 4; o; o >>= 1) {  // first maximum of the group's eight lanes: larger value, then smaller index
        const double ov = __shfl_xor_sync(0xffffffffu, maximum, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > maximum || (ov == maximum && oi < bi)) { maximum = ov; bi = oi; }
      }
      const bool rot = !done && maximum > DBL_EPSILON;
      const int p = rot ? (bi < pivotRow ? bi : pivotRow) : 0, q = rot ? (bi < pivotRow ? pivotRow : bi) : 1;
      if (!done) {
        pivotRow = (pivotRow + 1) % N;
        if (rot) zeroCount = 0;
        else if (++zeroCount == N) done = true;
      }
      if (__any_sync(0xffffffffu, rot)) {  // warp-uniform: an iteration in which no group rotates costs only its pivot scan
        const double Dpq = rot ? D[SIM_TRI(p, q)] : 1.0;
        const double theta = 0.5 * ((rot ? D[SIM_TRI(q, q)] : 0.0) - (rot ? D[SIM_TRI(p, p)] : 0.0)) / Dpq;
        double t = 1 / (fabs(theta) + sim_hypot(theta, 1));
        if (theta < 0) t = -t;
        const double c = 1 / sim_hypot(t, 1);
        const double s = c * t;
        const double tau = s / (1 + c);
        __syncwarp();
        if (rot) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = sl + 8 * u;
            if (r < N) {
              if (r < p) {
                const double Drp = D[SIM_TRI(r, p)], Drq = D[SIM_TRI(r, q)];
                D[SIM_TRI(r, p)] -= s * (Drq + tau * Drp);
                D[SIM_TRI(r, q)] += s * (Drp - tau * Drq);
              } else if (r > p && r < q) {
                const double Drp = D[SIM_TRI(p, r)], Drq = D[SIM_TRI(r, q)];
                D[SIM_TRI(p, r)] -= s * (Drq + tau * Drp);
                D[SIM_TRI(r, q)] += s * (Drp - tau * Drq);
              } else if (r > q) {
                const double Drp = D[SIM_TRI(p, r)], Drq = D[SIM_TRI(q, r)];
                D[SIM_TRI(p, r)] -= s * (Drq + tau * Drp);
                D[SIM_TRI(q, r)] += s * (Drp - tau * Drq);
              } else if (r == p) {
                D[SIM_TRI(p, p)] -= t * Dpq;
                D[SIM_TRI(q, q)] += t * Dpq;
                D[SIM_TRI(p, q)] = 0;
              }
            } else {
              const int rr = r - N;
              const double Vrp = Vt[p * N + rr], Vrq = Vt[q * N + rr];
              Vt[p * N + rr] -= s * (Vrq + tau * Vrp);
              Vt[q * N + rr] += s * (Vrp - tau * Vrq);
            }
          }
        }
        __syncwarp();
      }
    }
    if (a < n) {  // sample = SVec * sqrt(SVal) * normals + 0  (SIM2:19-31)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = sl + 8 * u;
        double acc = 0.0;
        for (int j = 0; j < N; ++j) acc += (Vt[j * N + i] * sqrt(D[SIM_TRI(j, j)])) * nrm[j];
        sc.samp[(size_t)a * 16 + i] = acc + 0.0;
      }
    }
    __syncwarp();
  }
}

#ifndef SIM_FILTER_CTAS
#define SIM_FILTER_CTAS 5  // x 4 warps, 11.3 KB of shared memory per warp
#endif
__global__ void __launch_bounds__(128, SIM_FILTER_CTAS) k_sim_filter(int n, const LqroSimModel* __restrict__ model, double* __restrict__ x,
                                                       double* __restrict__ Rot, double* __restrict__ Pcov, double* __restrict__ xTrue,
                                                       double* __restrict__ RotTrue, const double* __restrict__ pGoal,
                                                       const double* __restrict__ normals, LqroSimScratch sc, double* __restrict__ vGoal) {
  extern __shared__ __align__(16) unsigned char sim_smem[];
  const LqroSimModel& md = *model;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  LqroSimWsC& w = *reinterpret_cast<LqroSimWsC*>(sim_smem + (size_t)warp * sizeof(LqroSimWsC));
  for (int a = blockIdx.x * wpb + warp; a < n; a += gridDim.x * wpb) {
    for (int e = lane; e < 256; e += 32) w.P[e] = Pcov[(size_t)a * 256 + e];
    if (lane < 16) { w.x[lane] = x[(size_t)a * 16 + lane]; w.xT[lane] = xTrue[(size_t)a * 16 + lane]; }
    if (lane < 9) { w.R[lane] = Rot[(size_t)a * 9 + lane]; w.RT[lane] = RotTrue[(size_t)a * 9 + lane]; }
    if (lane < 6) w.nrm[lane] = normals[(size_t)a * 22 + 16 + lane];
    __syncwarp();
    double pg[3], vG[3];
    for (int i = 0; i < 3; ++i) pg[i] = pGoal[(size_t)a * 3 + i];
    sim_phase_filter<SimWarp>(w, sc.Ae + (size_t)a * 256, sc.MM + (size_t)(2 * a + 1) * 256, sc.v1 + (size_t)(2 * a + 1) * 16,
                              sc.v1 + (size_t)(2 * a) * 16, sc.samp + (size_t)a * 16, pg, vG, md);
    for (int e = lane; e < 256; e += 32) Pcov[(size_t)a * 256 + e] = w.P[e];
    if (lane < 16) { x[(size_t)a * 16 + lane] = w.x[lane]; xTrue[(size_t)a * 16 + lane] = w.xT[lane]; }
    if (lane < 9) { Rot[(size_t)a * 9 + lane] = w.R[lane]; RotTrue[(size_t)a * 9 + lane] = w.RT[lane]; }
    if (lane == 0) { vGoal[(size_t)a * 3] = vG[0]; vGoal[(size_t)a * 3 + 1] = vG[1]; vGoal[(size_t)a * 3 + 2] = vG[2]; }
    __syncwarp();
  }
}
#endif
