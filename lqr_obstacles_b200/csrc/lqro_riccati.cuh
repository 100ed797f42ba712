// lqro_riccati.cuh -- batched discrete-time Riccati solve for the 16-state quadrotor
// (controlMatrices MAIN:512-543): velocity LQR (S, T recursion -> L, E) followed by the position LQR
// on the closed loop (Stilde, Ttilde recursion with cross term -> Lh, Eh).
//
// Mapping: one warp per problem, all 16x16 / 16xm blocks resident in shared memory, each lane owning
// output elements of every small product; the m x m (4x4, 3x3) inverse is done by lane 0.  These
// matrices are far too small for tensor cores.  The recursion is a contraction run to its fixed point
// (300 iterations are converged to 4e-17 relative), so this kernel may use FMA: gains agree
// with the reference to ~1e-14 relative.
#pragma once
#include "lqro_common.cuh"

// C (n x p) = op(A) * B,  op(A) = A (n x m) or A^T (A stored m x n);  optional  C = D +/- op(A)*B
template <bool TA>
__device__ __forceinline__ void wmm(double* C, const double* A, const double* B, int n, int m, int p, const double* D = nullptr,
                                    double sign = 1.0) {
  const int lane = lane_id();
  for (int e = lane; e < n * p; e += 32) {
    const int i = e / p, j = e - i * p;
    double acc = 0.0;
    for (int k = 0; k < m; ++k) acc = fma(TA ? A[k * n + i] : A[i * m + k], B[k * p + j], acc);
    C[e] = D ? D[e] + sign * acc : sign * acc;
  }
  __syncwarp();
}

// in-place inverse of an m x m matrix (m <= 4), Gauss-Jordan with partial pivoting, lane 0
__device__ __forceinline__ void small_inverse(double* M, double* Minv, int m) {
  if (lane_id() == 0) {
    double a[16], b[16];
    for (int i = 0; i < m * m; ++i) { a[i] = M[i]; b[i] = 0.0; }
    for (int i = 0; i < m; ++i) b[i * m + i] = 1.0;
    for (int c = 0; c < m; ++c) {
      int piv = c;
      double best = fabs(a[c * m + c]);
      for (int r = c + 1; r < m; ++r)
        if (fabs(a[r * m + c]) > best) { best = fabs(a[r * m + c]); piv = r; }
      if (piv != c)
        for (int k = 0; k < m; ++k) {
          double t = a[c * m + k]; a[c * m + k] = a[piv * m + k]; a[piv * m + k] = t;
          t = b[c * m + k]; b[c * m + k] = b[piv * m + k]; b[piv * m + k] = t;
        }
      const double d = 1.0 / a[c * m + c];
      for (int k = 0; k < m; ++k) { a[c * m + k] *= d; b[c * m + k] *= d; }
      for (int r = 0; r < m; ++r) {
        if (r == c) continue;
        const double f = a[r * m + c];
        for (int k = 0; k < m; ++k) { a[r * m + k] -= f * a[c * m + k]; b[r * m + k] -= f * b[c * m + k]; }
      }
    }
    for (int i = 0; i < m * m; ++i) Minv[i] = b[i];
  }
  __syncwarp();
}

struct RiccatiWs {  // per-warp shared-memory workspace (doubles)
  double A[256], B[64], Q[256], S[256], S2[256], SA[256];
  double T[48], T2[48], Tq[48];
  double SB[64], AtSB[64], W[64], U[64], Pc[64], BtT[12];
  double Rm[16], M[16], Minv[16];
  double L[64], E[12];
};

// S <- Q + A'SA - (Pc' + A'SB) (Rm + B'SB)^-1 (Pc + B'SA),  T <- Tq + A'T - (Pc' + A'SB) (..)^-1 B'T
// and on exit  gain = -(..)^-1 (Pc + B'SA)  (m x 16),  gainE = -(..)^-1 B'T  (m x 3)
__device__ __forceinline__ void riccati_solve(RiccatiWs& w, int m, int iters, double* gain, double* gainE) {
  const int lane = lane_id();
  for (int it = 0; it <= iters; ++it) {
    wmm<false>(w.SB, w.S, w.B, 16, 16, m);              // S B
    wmm<true>(w.M, w.B, w.SB, m, 16, m, w.Rm, 1.0);     // Rm + B' S B
    small_inverse(w.M, w.Minv, m);
    wmm<false>(w.SA, w.S, w.A, 16, 16, 16);             // S A
    wmm<true>(w.U, w.B, w.SA, m, 16, 16, w.Pc, 1.0);    // Pc + B' S A     (m x 16)
    wmm<true>(w.BtT, w.B, w.T, m, 16, 3);               // B' T            (m x 3)
    if (it == iters) {                                   // gains from the converged S, T (MAIN:522-523, 542-543)
      wmm<false>(gain, w.Minv, w.U, m, m, 16, nullptr, -1.0);
      wmm<false>(gainE, w.Minv, w.BtT, m, m, 3, nullptr, -1.0);
      break;
    }
    wmm<true>(w.AtSB, w.A, w.SB, 16, 16, m);            // A' S B          (16 x m)
    for (int e = lane; e < 16 * m; e += 32) {            // + Pc'
      const int i = e / m, j = e - i * m;
      w.AtSB[e] += w.Pc[j * 16 + i];
    }
    __syncwarp();
    wmm<false>(w.W, w.AtSB, w.Minv, 16, m, m);          // (Pc' + A'SB) Minv
    wmm<true>(w.T2, w.A, w.T, 16, 16, 3, w.Tq, 1.0);    // Tq + A' T
    wmm<false>(w.T2, w.W, w.BtT, 16, m, 3, w.T2, -1.0); // - W B'T
    wmm<true>(w.S2, w.A, w.SA, 16, 16, 16, w.Q, 1.0);   // Q + A' S A
    wmm<false>(w.S2, w.W, w.U, 16, m, 16, w.S2, -1.0);  // - W (Pc + B'SA)
    for (int e = lane; e < 256; e += 32) w.S[e] = w.S2[e];
    for (int e = lane; e < 48; e += 32) w.T[e] = w.T2[e];
    __syncwarp();
  }
}

// grid: ceil(n_batch / warps_per_block) blocks; dynamic shared memory = warps_per_block * sizeof(RiccatiWs)
__global__ void k_riccati(int n_batch, const double* __restrict__ A, const double* __restrict__ B, double qv, double qp, double r,
                          int iters, double* __restrict__ Lout, double* __restrict__ Eout, double* __restrict__ Lhout,
                          double* __restrict__ Ehout) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int prob = blockIdx.x * (blockDim.x >> 5) + warp;
  if (prob >= n_batch) return;
  RiccatiWs& w = *reinterpret_cast<RiccatiWs*>(smem + (size_t)warp * ((sizeof(RiccatiWs) + 7) / 8));
  // ---- velocity LQR: Q = V'QvV, Tq = -V'Qv, Rm = r I4, Pc = 0, S0 = Q, T0 = Tq (MAIN:512-513)
  for (int e = lane; e < 256; e += 32) {
    w.A[e] = A[(size_t)prob * 256 + e];
    const int i = e >> 4, j = e & 15;
    const double q = (i == j && i >= 3 && i < 6) ? qv : 0.0;
    w.Q[e] = q; w.S[e] = q;
  }
  for (int e = lane; e < 64; e += 32) { w.B[e] = B[(size_t)prob * 64 + e]; w.Pc[e] = 0.0; }
  for (int e = lane; e < 48; e += 32) {
    const int i = e / 3, j = e - 3 * i;
    const double t = (i == j + 3) ? -qv : 0.0;
    w.Tq[e] = t; w.T[e] = t;
  }
  for (int e = lane; e < 16; e += 32) w.Rm[e] = ((e % 5) == 0) ? r : 0.0;
  __syncwarp();
  riccati_solve(w, 4, iters, w.L, w.E);
  for (int e = lane; e < 64; e += 32) Lout[(size_t)prob * 64 + e] = w.L[e];
  for (int e = lane; e < 12; e += 32) Eout[(size_t)prob * 12 + e] = w.E[e];
  __syncwarp();
  // ---- position LQR on the closed loop (MAIN:526-538):
  //   Qtilde = Pmap'QpPmap + 0.25 L'RL, Rtilde = 0.25 E'RE, Ptilde = 0.25 E'RL, Atilde = A + BL, Btilde = BE
  //   (scratch: SA <- Atilde, SB <- Btilde)
  wmm<false>(w.SA, w.B, w.L, 16, 4, 16, w.A, 1.0);
  wmm<false>(w.SB, w.B, w.E, 16, 4, 3);
  wmm<true>(w.S2, w.L, w.L, 16, 4, 16);   // L'L
  wmm<true>(w.M, w.E, w.E, 3, 4, 3);      // E'E
  wmm<true>(w.U, w.E, w.L, 3, 4, 16);     // E'L
  for (int e = lane; e < 256; e += 32) {
    const int i = e >> 4, j = e & 15;
    const double q = 0.25 * r * w.S2[e] + ((i == j && i < 3) ? qp : 0.0);
    w.Q[e] = q; w.S[e] = q;
    w.A[e] = w.SA[e];
  }
  __syncwarp();
  for (int e = lane; e < 48; e += 32) {
    w.B[e] = w.SB[e];
    w.Pc[e] = 0.25 * r * w.U[e];
    const int i = e / 3, j = e - 3 * i;
    const double t = (i == j) ? -qp : 0.0;
    w.Tq[e] = t; w.T[e] = t;
  }
  for (int e = lane; e < 9; e += 32) w.Rm[e] = 0.25 * r * w.M[e];
  __syncwarp();
  riccati_solve(w, 3, iters, w.L, w.E);
  for (int e = lane; e < 48; e += 32) Lhout[(size_t)prob * 48 + e] = w.L[e];
  for (int e = lane; e < 9; e += 32) Ehout[(size_t)prob * 9 + e] = w.E[e];
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
