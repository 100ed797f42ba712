// lqro_riccati.cuh -- batched discrete-time Riccati solve for the 16-state quadrotor
// (controlMatrices MAIN:512-543): velocity LQR (S, T recursion -> L, E) followed by the position LQR
// on the closed loop (Stilde, Ttilde recursion with cross term -> Lh, Eh).
//
//   S <- Q + A'SA - (Pc' + A'SB) (Rm + B'SB)^-1 (Pc + B'SA),   T <- Tq + A'T - (Pc' + A'SB) (..)^-1 B'T,    `iters` times,
//   gain = -(..)^-1 (Pc + B'SA),  gainE = -(..)^-1 B'T  from the last S, T.
// The count is part of the contract: the reference runs exactly 300 iterations and its position recursion is NOT at its
// fixed point by then (the iterate still moves by 2e-6 relative per iteration), so a doubling algorithm or an early exit
// would return other gains.  The recursions themselves are contractions, so this kernel may use FMA and its own summation
// order: gains agree with the reference to ~1e-14 relative (bar: 1e-9).
//
// MAPPING: ONE WARP PER PROBLEM, the recursion's state in registers.  Lane l = (row i = l / 2, half h = l % 2) owns the
// half rows S2[i][8h .. 8h+7] of every 16 x 16 result (8 independent accumulators per lane: the FMA chains of the products
// are pipelined, not latency-bound) and keeps, across all iterations, the full row i of S and the full row i of A' in
// registers.  What streams from shared memory is read as 16-byte vectors at two distinct addresses per instruction
// (broadcast): rows of A / S A, of B / S B / T.  The 16 x m blocks ride along in the same k loops, two columns per lane
// (m = 3 is padded to 4 with a zero column of B and a unit pivot).  Per iteration: phase 1  S A, S B (registers x shared
// A, B) -> shared; phase 2  Q + A'(S A), A'(S B), Tq + A'T (registers x shared), Pc + B'(S A) two entries per lane,
// Rm + B'SB and B'T on the two half warps, the m x m inverse by a Gauss-Jordan sweep over the
// 4 x 8 augmented matrix with ONE ELEMENT PER LANE (pivot row / column by shuffle; the matrix is symmetric positive
// definite, no pivoting); phase 3  W = (Pc' + A'SB) Minv, the rank-m corrections, and the lane pair swaps half rows by
// shuffle so that both hold the new row of S.  Two __syncwarp per iteration.  These matrices are far too small for
// tensor cores (north_star).
//
// DUAL COMPILATION: riccati_warp() also compiles as plain C++ on tests/hull_host/warp_emul.h (32 lockstep host threads),
// tests/test_riccati_host.py checks it against the oracle where no GPU exists.  liblqro.so never contains that build.
#pragma once
#ifndef LQRO_HULL_HOST_EMUL
#include "lqro_common.cuh"
#endif

struct RicD2 { double x, y; };
__device__ __forceinline__ RicD2 ric_ld2(const double* p) {
#ifdef LQRO_HULL_HOST_EMUL
  RicD2 v; v.x = p[0]; v.y = p[1]; return v;
#else
  const double2 t = *reinterpret_cast<const double2*>(p);
  RicD2 v; v.x = t.x; v.y = t.y; return v;
#endif
}
__device__ __forceinline__ void ric_st2(double* p, double a, double b) {
#ifdef LQRO_HULL_HOST_EMUL
  p[0] = a; p[1] = b;
#else
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
#endif
}

// 16 x 16 blocks that lanes write or read as their own half rows (S A, Q) are stored with the 16-byte vectors of a half row
// permuted by the row: vector q of row i sits at q ^ (i & 3).  Eight lanes of a quarter warp (rows i .. i+3, both halves) then hit
// eight different bank groups instead of two (the plain layout costs a 4-way conflict on every such access).
__device__ __forceinline__ int ric_swz(int row, int q) { return 2 * (q ^ (row & 3)); }
// 1 / x to full precision without the division routine: hardware seed + two Newton steps (the recursion needs no IEEE quotient)
__device__ __forceinline__ double ric_rcp(double x) {
#ifdef LQRO_HULL_HOST_EMUL
  return 1.0 / x;
#else
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(fma(-x, y, 1.0), y, y);
  y = fma(fma(-x, y, 1.0), y, y);
  y = fma(fma(-x, y, 1.0), y, y);
  return y;
#endif
}

struct __align__(16) RiccatiWs {  // per-warp shared-memory workspace (doubles); 16 x 4 blocks are [16][4], 4 x 16 blocks [4][16]
  double A[256], SA[256], Q[256];
  double B[64], SB[64], U[64], Pc[64], T[64], Tq[64];
  double Minv[16], BtT[16], Rm[16];
  double L[64], E[16];
};

// `iters` iterations on the state (S row in registers, w.T), then the gains into w.L (4 x 16) / w.E (4 x 4, column 3 unused).
// In: w.A, w.Q, w.B, w.Pc, w.Tq, w.Rm set, w.T = T0; S[] = row i of S0, At[] = row i of A' (column i of A).
__device__ __forceinline__ void riccati_iterate(RiccatiWs& w, double (&S)[16], const double (&At)[16], int iters) {
  const int lane = lane_id();
  const int i = lane >> 1, h = lane & 1, c0 = 8 * h;
  for (int it = 0;; ++it) {
    // ---- phase 1: S A (half row) and S B (two columns)
    {
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      double sb0 = 0, sb1 = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double s = S[k];
        const double* ar = w.A + k * 16 + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const RicD2 a = ric_ld2(ar + 2 * q);
          acc[2 * q] = fma(s, a.x, acc[2 * q]);
          acc[2 * q + 1] = fma(s, a.y, acc[2 * q + 1]);
        }
        const RicD2 b = ric_ld2(w.B + k * 4 + 2 * h);
        sb0 = fma(s, b.x, sb0);
        sb1 = fma(s, b.y, sb1);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) ric_st2(w.SA + i * 16 + c0 + ric_swz(i, q), acc[2 * q], acc[2 * q + 1]);
      ric_st2(w.SB + i * 4 + 2 * h, sb0, sb1);
    }
    __syncwarp();
    // ---- phase 2: Q + A'(S A), A'(S B), Tq + A'T; Rm + B'SB | B'T; inverse
    double acc[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const RicD2 v = ric_ld2(w.Q + i * 16 + c0 + ric_swz(i, q)); acc[2 * q] = v.x; acc[2 * q + 1] = v.y; }
    double up0 = 0, up1 = 0, t0, t1;
    { const RicD2 v = ric_ld2(w.Tq + i * 4 + 2 * h); t0 = v.x; t1 = v.y; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const double a = At[k];
      const double* sr = w.SA + k * 16 + c0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const RicD2 v = ric_ld2(sr + ric_swz(k, q));
        acc[2 * q] = fma(a, v.x, acc[2 * q]);
        acc[2 * q + 1] = fma(a, v.y, acc[2 * q + 1]);
      }
      const RicD2 b = ric_ld2(w.SB + k * 4 + 2 * h);
      up0 = fma(a, b.x, up0);
      up1 = fma(a, b.y, up1);
      const RicD2 t = ric_ld2(w.T + k * 4 + 2 * h);
      t0 = fma(a, t.x, t0);
      t1 = fma(a, t.y, t1);
    }
    up0 += w.Pc[(2 * h) * 16 + i];       // (Pc' + A'SB)[i][2h]
    up1 += w.Pc[(2 * h + 1) * 16 + i];
    // U = Pc + B'(S A), two entries per lane.  Not taken as the transpose of A'SB + Pc': the computed S is symmetric only up to
    // rounding, and a recursion that assumes symmetry amplifies the antisymmetric part (A has eigenvalues on the unit circle)
    {
      const int a = lane >> 3, jh = (lane & 7) >> 2, jq = lane & 3, j = 2 * (lane & 7);
      RicD2 u = ric_ld2(w.Pc + a * 16 + j);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double b = w.B[k * 4 + a];
        const RicD2 v = ric_ld2(w.SA + k * 16 + 8 * jh + ric_swz(k, jq));
        u.x = fma(b, v.x, u.x);
        u.y = fma(b, v.y, u.y);
      }
      ric_st2(w.U + a * 16 + j, u.x, u.y);
    }
    // lanes 0..15: (Rm + B'SB)[a][b]; lanes 16..31: (B'T)[a][c]   (a = (lane & 15) / 4, b or c = lane & 3)
    double mval;
    {
      const int a = (lane & 15) >> 2, b = lane & 3;
      const double* X = lane < 16 ? w.SB : w.T;
      mval = lane < 16 ? w.Rm[lane] : 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) mval = fma(w.B[k * 4 + a], X[k * 4 + b], mval);
      if (lane >= 16) w.BtT[lane - 16] = mval;
    }
    // Gauss-Jordan on [M | I], element (r, c) of the 4 x 8 augmented matrix on lane r * 8 + c
    {
      const int r = lane >> 3, c = lane & 7;
      double e = __shfl_sync(LQRO_FULL, mval, c < 4 ? r * 4 + c : lane);
      if (c >= 4) e = (c - 4 == r) ? 1.0 : 0.0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const double piv = __shfl_sync(LQRO_FULL, e, p * 8 + p);
        const double rowp = __shfl_sync(LQRO_FULL, e, p * 8 + c);
        const double f = __shfl_sync(LQRO_FULL, e, r * 8 + p);
        const double q = rowp * ric_rcp(piv);
        e = (r == p) ? q : fma(-f, q, e);
      }
      if (c >= 4) w.Minv[r * 4 + (c - 4)] = e;
    }
    __syncwarp();
    if (it == iters) {  // gains from the last S, T (MAIN:522-523, 542-543)
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int idx = 2 * lane + z, a = idx >> 4, j = idx & 15;
        double g = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) g = fma(w.Minv[a * 4 + b], w.U[b * 16 + j], g);
        w.L[idx] = -g;
      }
      if (lane < 16) {
        const int a = lane >> 2, c = lane & 3;
        double g = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) g = fma(w.Minv[a * 4 + b], w.BtT[b * 4 + c], g);
        w.E[lane] = -g;
      }
      __syncwarp();
      return;
    }
    // ---- phase 3: W = (Pc' + A'SB) Minv (row i), rank-m corrections, new row of S
    {
      const double o0 = __shfl_xor_sync(LQRO_FULL, up0, 1), o1 = __shfl_xor_sync(LQRO_FULL, up1, 1);
      const double u0 = h ? o0 : up0, u1 = h ? o1 : up1, u2 = h ? up0 : o0, u3 = h ? up1 : o1;
      double W[4];
#pragma unroll
      for (int a = 0; a < 4; a += 2) {
        const RicD2 m0 = ric_ld2(w.Minv + 0 + a), m1 = ric_ld2(w.Minv + 4 + a), m2 = ric_ld2(w.Minv + 8 + a), m3 = ric_ld2(w.Minv + 12 + a);
        W[a] = fma(u3, m3.x, fma(u2, m2.x, fma(u1, m1.x, u0 * m0.x)));
        W[a + 1] = fma(u3, m3.y, fma(u2, m2.y, fma(u1, m1.y, u0 * m0.y)));
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double wa = -W[a];
        const double* ur = w.U + a * 16 + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const RicD2 v = ric_ld2(ur + 2 * q);
          acc[2 * q] = fma(wa, v.x, acc[2 * q]);
          acc[2 * q + 1] = fma(wa, v.y, acc[2 * q + 1]);
        }
        const RicD2 bt = ric_ld2(w.BtT + a * 4 + 2 * h);
        t0 = fma(wa, bt.x, t0);
        t1 = fma(wa, bt.y, t1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double o = __shfl_xor_sync(LQRO_FULL, acc[j], 1);
        S[j] = h ? o : acc[j];
        S[8 + j] = h ? acc[j] : o;
      }
      ric_st2(w.T + i * 4 + 2 * h, t0, t1);  // read again in phase 2 of the next iteration, behind the phase-1 barrier
    }
  }
}

// One problem on one warp.  A (16 x 16), B (16 x 4) row major in global memory; results L (4 x 16), E (4 x 3), Lh (3 x 16), Eh (3 x 3).
__device__ __forceinline__ void riccati_warp(RiccatiWs& w, const double* __restrict__ A, const double* __restrict__ B, double qv, double qp,
                                             double r, int iters, double* __restrict__ Lout, double* __restrict__ Eout,
                                             double* __restrict__ Lhout, double* __restrict__ Ehout) {
  const int lane = lane_id();
  const int i = lane >> 1;
  // ---- velocity LQR: Q = V'QvV, Tq = -V'Qv, Rm = r I4, Pc = 0, S0 = Q, T0 = Tq (MAIN:512-513)
  for (int e = lane; e < 256; e += 32) {
    w.A[e] = A[e];
    const int a = e >> 4, b = e & 15;
    w.Q[a * 16 + (b & 8) + ric_swz(a, (b >> 1) & 3) + (b & 1)] = (a == b && a >= 3 && a < 6) ? qv : 0.0;
  }
  for (int e = lane; e < 64; e += 32) {
    w.B[e] = B[e]; w.Pc[e] = 0.0;
    const int a = e >> 2, c = e & 3;
    const double t = (c < 3 && a == c + 3) ? -qv : 0.0;
    w.Tq[e] = t; w.T[e] = t;
  }
  if (lane < 16) w.Rm[lane] = ((lane % 5) == 0) ? r : 0.0;
  __syncwarp();
  double S[16], At[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { S[k] = w.Q[i * 16 + (k & 8) + ric_swz(i, (k >> 1) & 3) + (k & 1)]; At[k] = w.A[k * 16 + i]; }
  riccati_iterate(w, S, At, iters);
  for (int e = lane; e < 64; e += 32) Lout[e] = w.L[e];
  if (lane < 12) Eout[lane] = w.E[(lane / 3) * 4 + lane % 3];
  // ---- position LQR on the closed loop (MAIN:526-538):
  //   Atilde = A + B L, Btilde = B E, Qtilde = Pmap'QpPmap + 0.25 r L'L, Rtilde = 0.25 r E'E, Ptilde = 0.25 r E'L
  //   (one-off: plain loops; SA holds Atilde, SB Btilde until they are copied over A, B)
  for (int e = lane; e < 256; e += 32) {
    const int a = e >> 4, b = e & 15;
    double at = w.A[e], ll = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { at = fma(w.B[a * 4 + k], w.L[k * 16 + b], at); ll = fma(w.L[k * 16 + a], w.L[k * 16 + b], ll); }
    w.SA[e] = at;
    w.Q[a * 16 + (b & 8) + ric_swz(a, (b >> 1) & 3) + (b & 1)] = 0.25 * r * ll + ((a == b && a < 3) ? qp : 0.0);
  }
  for (int e = lane; e < 64; e += 32) {
    const int a = e >> 2, c = e & 3;  // Btilde[a][c], c < 3
    double bt = 0;
    if (c < 3) {
#pragma unroll
      for (int k = 0; k < 4; ++k) bt = fma(w.B[a * 4 + k], w.E[k * 4 + c], bt);
    }
    w.SB[e] = bt;
    const int pa = e >> 4, pj = e & 15;  // Ptilde[pa][pj], pa < 3
    double pc = 0;
    if (pa < 3) {
#pragma unroll
      for (int k = 0; k < 4; ++k) pc = fma(w.E[k * 4 + pa], w.L[k * 16 + pj], pc);
    }
    w.Pc[e] = 0.25 * r * pc;
    const double t = (c < 3 && a == c) ? -qp : 0.0;
    w.Tq[e] = t; w.T[e] = t;
  }
  if (lane < 16) {
    const int a = lane >> 2, b = lane & 3;
    double ee = 0;
    if (a < 3 && b < 3) {
#pragma unroll
      for (int k = 0; k < 4; ++k) ee = fma(w.E[k * 4 + a], w.E[k * 4 + b], ee);
    }
    w.Rm[lane] = (a < 3 && b < 3) ? 0.25 * r * ee : (a == b ? 1.0 : 0.0);  // unit pivot for the padded control
  }
  __syncwarp();
  for (int e = lane; e < 256; e += 32) w.A[e] = w.SA[e];
  for (int e = lane; e < 64; e += 32) w.B[e] = w.SB[e];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) { S[k] = w.Q[i * 16 + (k & 8) + ric_swz(i, (k >> 1) & 3) + (k & 1)]; At[k] = w.A[k * 16 + i]; }
  riccati_iterate(w, S, At, iters);
  for (int e = lane; e < 48; e += 32) Lhout[e] = w.L[e];
  if (lane < 9) Ehout[lane] = w.E[(lane / 3) * 4 + lane % 3];
  __syncwarp();
}

#ifndef LQRO_HULL_HOST_EMUL
// grid: ceil(n_batch / warps_per_block) blocks; dynamic shared memory = warps_per_block * sizeof(RiccatiWs)
__global__ void __launch_bounds__(128) k_riccati(int n_batch, const double* __restrict__ A, const double* __restrict__ B, double qv, double qp, double r,
                                                 int iters, double* __restrict__ Lout, double* __restrict__ Eout, double* __restrict__ Lhout,
                                                 double* __restrict__ Ehout) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  const int prob = blockIdx.x * (blockDim.x >> 5) + warp;
  if (prob >= n_batch) return;
  RiccatiWs& w = *reinterpret_cast<RiccatiWs*>(smem + (size_t)warp * (sizeof(RiccatiWs) / 8));
  riccati_warp(w, A + (size_t)prob * 256, B + (size_t)prob * 64, qv, qp, r, iters, Lout + (size_t)prob * 64, Eout + (size_t)prob * 12,
               Lhout + (size_t)prob * 48, Ehout + (size_t)prob * 9);
}

// ------------------------------------------------------------------------------------------------
// Peak-rate micro-benchmarks (roofline denominators not in MEASURED_PEAKS.json): 8 independent FMA
// chains per thread, enough threads to fill every SM.
__global__ void k_peak_fp64(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_peak_fp32(float* out, int iters) {
  float a0 = threadIdx.x * 1e-6f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = 1.00001f, c = 1e-5f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
#endif
