// lqro_pair.cuh -- horizon tables and the per-pair LQR-obstacle kernel.
//
//   k_horizon_tables : findFG x H (MAIN:653-662) + the pair-independent half of createObstacle
//                      (MAIN:684-685), once per dynamics class.
//   k_pairs          : for every (agent, neighbour slot): createObstacle x H (MAIN:683-688),
//                      findReachableObstacle (MAIN:691-706), pointInHull/run_gjk (MAIN:712-758) and,
//                      for the outside case, createHalfPlanes (MAIN:1085-1098).  Inside pairs are queued
//                      for k_hull (lqro_hull.cuh).
//
// Data layout in HBM: agent state AoS x[N][16] FP64 (one 128-byte line per agent: a pair costs one
// line for the neighbour), neighbour ids nbr[n_local][K] int32, outputs planes[n_local][K][6] FP32,
// status/nreach[n_local][K] int32, dist[n_local][K] FP64.  Tables live in shared memory.
#pragma once
#include "lqro_common.cuh"
#include "lqro_gjk.cuh"

// ------------------------------------------------------------------------------------------------
// 3x3 inverse exactly as the reference's operator! (MAT:603-671): Gauss-Jordan, full pivoting with
// row/column permutation vectors, the same update order.
__device__ inline void inverse3_ref(const double* q, double* out) {
  const int n = 3;
  double m[9], inv[9];
  int rp[3], cp[3];
  for (int i = 0; i < 9; ++i) { m[i] = q[i]; inv[i] = 0.0; }
  for (int i = 0; i < n; ++i) { inv[i * n + i] = 1.0; rp[i] = i; cp[i] = i; }
  for (int k = 0; k < n; ++k) {
    double maximum = 0;
    int mr = k, mc = k;
    for (int i = k; i < n; ++i)
      for (int j = k; j < n; ++j) {
        const double a = fabs(m[rp[i] * n + cp[j]]);
        if (a > maximum) { maximum = a; mr = i; mc = j; }
      }
    int t = rp[k]; rp[k] = rp[mr]; rp[mr] = t;
    t = cp[k]; cp[k] = cp[mc]; cp[mc] = t;
    for (int i = k + 1; i < n; ++i) {
      const double factor = m[rp[i] * n + cp[k]] / m[rp[k] * n + cp[k]];
      for (int j = k + 1; j < n; ++j) m[rp[i] * n + cp[j]] -= factor * m[rp[k] * n + cp[j]];
      for (int j = 0; j < k; ++j) inv[rp[i] * n + rp[j]] -= factor * inv[rp[k] * n + rp[j]];
      inv[rp[i] * n + rp[k]] = -factor;
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    const double quotient = m[rp[k] * n + cp[k]];
    for (int j = 0; j < n; ++j) inv[rp[k] * n + j] /= quotient;
    for (int i = 0; i < k; ++i) {
      const double factor = m[rp[i] * n + cp[k]];
      for (int j = 0; j < n; ++j) inv[rp[i] * n + j] -= factor * inv[rp[k] * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[cp[i] * n + j] = inv[rp[i] * n + j];
}

#ifndef LQRO_HULL_HOST_EMUL
// One block of 256 threads; thread t owns element (t/16, t%16) of the 16x16 products.  Each output
// element is one left-to-right sum (MAT:218-231), so the thread-parallel product is bit-identical to
// the reference's triple loop.
//   out: NFt [16][3H] (transposed), NF [H][3][16] (reference order, for the parity API), T [H][9].
__global__ void __launch_bounds__(256) k_horizon_tables(const double* __restrict__ A, const double* __restrict__ B,
                                                        const double* __restrict__ L, const double* __restrict__ E, int H,
                                                        double* __restrict__ NFt, double* __restrict__ NF,
                                                        double* __restrict__ T, double* __restrict__ CG) {
  __shared__ double sAt[256], sBt[48], sF[256], sG[48], sF2[256], sG2[48];
  const int t = threadIdx.x;
  const int i = t >> 4, j = t & 15;
  {
    double bl = 0;  // (B*L)(i,j)
    for (int k = 0; k < 4; ++k) bl += B[i * 4 + k] * L[k * 16 + j];
    sAt[t] = A[t] + bl;  // Atilde = A + (B*L), MAIN:655
    sF[t] = (i == j) ? 1.0 : 0.0;
  }
  if (t < 48) {
    const int bi = t / 3, bj = t % 3;
    double be = 0;  // Btilde = B*E, MAIN:658
    for (int k = 0; k < 4; ++k) be += B[bi * 4 + k] * E[k * 3 + bj];
    sBt[t] = be;
    sG[t] = 0.0;
  }
  __syncthreads();
  for (int k = 0; k < H; ++k) {
    {
      double f = 0;  // F = Atilde*F, MAIN:660
      for (int m = 0; m < 16; ++m) f += sAt[i * 16 + m] * sF[m * 16 + j];
      sF2[t] = f;
    }
    if (t < 48) {
      const int gi = t / 3, gj = t % 3;
      double g = 0;  // G = Atilde*G + Btilde, MAIN:661
      for (int m = 0; m < 16; ++m) g += sAt[gi * 16 + m] * sG[m * 3 + gj];
      sG2[t] = g + sBt[t];
    }
    __syncthreads();
    sF[t] = sF2[t];
    if (t < 48) sG[t] = sG2[t];
    __syncthreads();
    if (t < 48) {  // rows 0..2 of -F: (-P)*F with P = [I3 0], MAIN:685
      const int c = t >> 4, m = t & 15;
      const double v = -sF[c * 16 + m];
      NF[(k * 3 + c) * 16 + m] = v;
      NFt[m * (3 * H) + k * 3 + c] = v;
    }
    if (t == 0) {  // Transform = !(P*G), MAIN:684
      double cg[9];
      for (int e = 0; e < 9; ++e) { cg[e] = sG[e]; CG[9 * k + e] = cg[e]; }  // rows 0..2 of G (16x3 row major)
      inverse3_ref(cg, T + 9 * k);
    }
    __syncthreads();
  }
}

#endif  // LQRO_HULL_HOST_EMUL

// ------------------------------------------------------------------------------------------------
struct PairOutputs {
  float* planes;   // [n_local*K][6]
  int* status;     // [n_local*K]
  int* nreach;     // [n_local*K]
  double* dist;    // [n_local*K]
  int* nfacets;    // [n_local*K]
};

struct HullQueue {
  int* items;      // pair ids of the inside case
  int* count;      // number appended (may exceed capacity: overflow)
  int capacity;
};

// shared-memory carve-up of one pair warp
struct PairWarpSmem {
  int* klist;   // [H]    horizon steps that survive the pruning test
  double* tr;   // [3H]   Translate_k, MAIN:685
  double* dx;   // [16]   x_i - x_j
  double* xi;   // [16]
  double* px;   // [HP]   reachable cloud, SoA
  double* py;
  double* pz;
};

#define LQRO_PRN 8  // doubles of pruning constants per horizon step
__host__ __device__ inline size_t pair_block_table_doubles(int H, int P) { return (size_t)16 * 3 * H + 9 * H + 3 * P + 9 * H + LQRO_PRN * H; }
__host__ __device__ inline size_t pair_warp_doubles(int H, int P) {
  return (size_t)3 * H + 32 + 3 * (size_t)H * P + (H + 1) / 2 + 1;
}

// exact reachable test, MAIN:703: pow(a,2)/pow(r,2) + pow(b,2)/pow(r,2) + pow(c,2)/pow(r,2) < 1.0
__device__ __noinline__ bool reach_exact(double a, double b, double c, double r2) {  // out of line: the fast test below almost always decides
  return (a * a) / r2 + (b * b) / r2 + (c * c) / r2 < 1.0;
}
// Same decision without the three divisions unless the sum is within 1e-9 of the threshold: the
// fused estimate differs from the reference's rounded sum by < 2e-15 relative, so outside the band
// both sides of the comparison agree; inside it the exact sequence is evaluated.
__device__ __forceinline__ bool reach_test(double a, double b, double c, double r2, double inv_r2) {
  const double est = fma(a, a, fma(b, b, c * c)) * inv_r2;
  if (fabs(est - 1.0) > 1.0e-9) return est < 1.0;
  return reach_exact(a, b, c, r2);
}

// The same rule with the band on the sum of squares itself (r2lo = r2 (1 - 1e-9), r2hi = r2 (1 + 1e-9)): the fused sum is
// within 4e-16 relative of the exact a^2 + b^2 + c^2 and the reference's rounded quotient sum within 5e-16 of exact / r2, so
// below r2lo the reference's sum is < 1 and above r2hi it is > 1; in between (and for NaN) the exact sequence decides.
__device__ __forceinline__ bool reach_test_band(double a, double b, double c, double r2, double r2lo, double r2hi) {
  const double ss = fma(a, a, fma(b, b, c * c));
  if (ss < r2lo) return true;
  if (ss > r2hi) return false;
  return reach_exact(a, b, c, r2);
}

// Translate_k = -(C F_k) (x_i - x_j) for all k: 3H independent 16-term left-to-right sums (MAT:218-231).
__device__ __forceinline__ void warp_translate(const double* __restrict__ sNFt, const double* __restrict__ dx, double* tr, int H) {
  const int lane = lane_id();
  const int n3 = 3 * H;
  for (int e = lane; e < n3; e += 32) {
    double t = 0;
#pragma unroll
    for (int m = 0; m < 16; ++m) t += sNFt[m * n3 + e] * dx[m];
    tr[e] = t;
  }
}

// candidate point (k, s): Transform_k * (p_s + Translate_k), MAIN:687
__device__ __forceinline__ void candidate_point(const double* __restrict__ sT, const double* __restrict__ sphx,
                                                const double* __restrict__ sphy, const double* __restrict__ sphz,
                                                const double* __restrict__ tr, int k, int s, double& q0, double& q1, double& q2) {
  const double* Tk = sT + 9 * k;
  const double v0 = sphx[s] + tr[3 * k], v1 = sphy[s] + tr[3 * k + 1], v2 = sphz[s] + tr[3 * k + 2];
  q0 = Tk[0] * v0 + Tk[1] * v1 + Tk[2] * v2;
  q1 = Tk[3] * v0 + Tk[4] * v1 + Tk[5] * v2;
  q2 = Tk[6] * v0 + Tk[7] * v1 + Tk[8] * v2;
}

// Block-level tables in shared memory.
struct PairTables {
  const double *NFt, *T, *sphx, *sphy, *sphz, *CG, *prn;
};

// EXACT PRUNING of whole horizon steps.  A candidate of step k is q = T_k u with u = p_s + tr_k, and
// q - v_rel = T_k (u - c_k) (up to the residual of the computed inverse), c_k = (C G_k) v_rel.  Every sample
// p_s lies in the box [-a, a] (a = per-axis max |p_s|), so u - c_k lies in the box w_k + [-a, a],
// w_k = tr_k - c_k, whose points have |coordinate i| >= g_i = max(0, |w_i| - a_i) and norm <= U = |w| + |a|.
// Two lower bounds of |q - v_rel| over that box:
//   (ball)     |T z| >= |z| / sigma_max(C G_k)            >= |g| / sigma_max
//   (diagonal) |T z| >= |diag(T) z| - |offdiag(T)|_F |z|  >= sqrt(sum_i T_ii^2 g_i^2) - |offdiag(T)|_F U
// (the second is the sharp one for this model: C G_k is diagonal up to round-off and 3-200x stiffer in x, y
// than in z).  The step is dropped when either bound exceeds
//     max_speed (1 + 1e-6) + 1e-11 (|T_k|_F (U + |c_k|_1 + 1) + kappa_k |v_rel|_1):
// the second term over-estimates (by > 10^3) the rounding error of the reference's own evaluation of q
// (<= ~2^-50 |T_k|_F |u| per component), of c_k, and the residual |T_k C G_k - I| |v_rel| of the computed
// inverse.  Then every exact AND every rounded |q - v_rel| exceeds max_speed, the reference's test
// `sum((q-c)^2/r^2) < 1` (MAIN:703) is false for all P samples, and step k adds nothing to the cloud.  The
// surviving steps are evaluated with the reference's arithmetic, so the cloud is identical, bit for bit
// (tests: GPU step with and without pruning; CPU check of the rule against the oracle).
// tr0..2: Translate_k as computed by the caller; tr_err: bound on |tr - (reference's rounded Translate_k)|_1
// (0 when tr is the reference-order value itself).
__device__ __forceinline__ bool step_prunable_v(const LqroParams& p, const PairTables& t, int k, double tr0, double tr1, double tr2,
                                                double tr_err, double vr0, double vr1, double vr2) {
  const double* G = t.CG + 9 * k;
  const double* c = t.prn + LQRO_PRN * k;
  const double c0 = fma(G[0], vr0, fma(G[1], vr1, G[2] * vr2));
  const double c1 = fma(G[3], vr0, fma(G[4], vr1, G[5] * vr2));
  const double c2 = fma(G[6], vr0, fma(G[7], vr1, G[8] * vr2));
  const double w0 = tr0 - c0, w1 = tr1 - c1, w2 = tr2 - c2;
  const double g0 = fmax(0.0, fabs(w0) - p.box[0] - tr_err), g1 = fmax(0.0, fabs(w1) - p.box[1] - tr_err),
               g2 = fmax(0.0, fabs(w2) - p.box[2] - tr_err);
  const double U = fabs(w0) + fabs(w1) + fabs(w2) + p.box[0] + p.box[1] + p.box[2] + 3.0 * tr_err;  // >= |z| for every z in the box
  const double c1n = fabs(c0) + fabs(c1) + fabs(c2);
  const double v1n = fabs(vr0) + fabs(vr1) + fabs(vr2);
  const double need = p.max_speed * (1.0 + 1.0e-6) + 1.0e-11 * (c[6] * (U + c1n + 1.0) + c[1] * v1n);
  // (ball) in position space: |g| > sigma_max * need   (compared as squares, both sides >= 0)
  const double gap2 = fma(g0, g0, fma(g1, g1, g2 * g2));
  const double rb = c[0] * need;
  if (gap2 * (1.0 - 1.0e-12) > rb * rb) return true;
  // (diagonal) in velocity space: sqrt(sum (T_ii g_i)^2) > need + |offdiag T|_F U
  const double d0 = c[2] * g0, d1 = c[3] * g1, d2 = c[4] * g2;
  const double rhs = need + c[5] * U;
  return fma(d0, d0, fma(d1, d1, d2 * d2)) * (1.0 - 1.0e-12) > rhs * rhs;
}
__device__ __forceinline__ bool step_prunable(const LqroParams& p, const PairTables& t, const double* tr, int k, double vr0, double vr1,
                                              double vr2) {
  return step_prunable_v(p, t, k, tr[3 * k], tr[3 * k + 1], tr[3 * k + 2], 0.0, vr0, vr1, vr2);
}

// Build the reachable cloud of one pair in shared memory (order preserved), return its size.
// Requires w.dx (and so v_rel = dx[3..5]) to be set; fills w.tr.
__device__ __forceinline__ int warp_reachable_cloud(const LqroParams& p, const PairTables& t, PairWarpSmem& w) {
  const int lane = lane_id();
  warp_translate(t.NFt, w.dx, w.tr, p.H);
  __syncwarp();
  const double vr0 = w.dx[3], vr1 = w.dx[4], vr2 = w.dx[5];
  // surviving horizon steps, ascending (order of the cloud is k-major, MAIN:1383-1389)
  int nk = 0;
  for (int kb = 0; kb < p.H; kb += 32) {
    const int k = kb + lane;
    bool keep = false;
    if (k < p.H) keep = !(p.prune && step_prunable(p, t, w.tr, k, vr0, vr1, vr2));
    const unsigned mask = __ballot_sync(LQRO_FULL, keep);
    if (keep) w.klist[nk + __popc(mask & ((1u << lane) - 1u))] = k;
    nk += __popc(mask);
  }
  __syncwarp();
  int count = 0;
  const int total = nk * p.P;
  int ki = 0, s = lane;
  while (s >= p.P) { s -= p.P; ++ki; }
  for (int base = 0; base < total; base += 32) {
    const int m = base + lane;
    bool pass = false;
    double q0 = 0, q1 = 0, q2 = 0;
    if (m < total) {
      candidate_point(t.T, t.sphx, t.sphy, t.sphz, w.tr, w.klist[ki], s, q0, q1, q2);
      pass = reach_test(q0 - vr0, q1 - vr1, q2 - vr2, p.r2, p.inv_r2);
    }
    const unsigned mask = __ballot_sync(LQRO_FULL, pass);
    if (pass) {
      const int pos = count + __popc(mask & ((1u << lane) - 1u));
      w.px[pos] = q0; w.py[pos] = q1; w.pz[pos] = q2;
    }
    count += __popc(mask);
    s += 32;
    while (s >= p.P) { s -= p.P; ++ki; }
  }
  __syncwarp();
  return count;
}

__device__ __forceinline__ void carve_pair_smem(double* base, int warp, int H, int P, PairTables& t, PairWarpSmem& w) {
  t.NFt = base;
  t.T = base + 16 * 3 * H;
  t.sphx = t.T + 9 * H;
  t.sphy = t.sphx + P;
  t.sphz = t.sphy + P;
  t.CG = t.sphz + P;
  t.prn = t.CG + 9 * H;
  double* wb = base + pair_block_table_doubles(H, P) + (size_t)warp * pair_warp_doubles(H, P);
  w.tr = wb;
  w.dx = wb + 3 * H;
  w.xi = w.dx + 16;
  w.px = w.xi + 16;
  w.py = w.px + (size_t)H * P;
  w.pz = w.py + (size_t)H * P;
  w.klist = reinterpret_cast<int*>(w.pz + (size_t)H * P);
}

#ifndef LQRO_HULL_HOST_EMUL
__device__ __forceinline__ void load_block_tables(double* base, const LqroTables& tb, int H, int P) {
  const int nNF = 16 * 3 * H, nT = 9 * H;
  for (int e = threadIdx.x; e < nNF; e += blockDim.x) base[e] = tb.NFt[e];
  for (int e = threadIdx.x; e < nT; e += blockDim.x) base[nNF + e] = tb.T[e];
  double* sph = base + nNF + nT;
  for (int e = threadIdx.x; e < P; e += blockDim.x) {  // AoS [P][3] -> SoA
    sph[e] = tb.sph[3 * e];
    sph[P + e] = tb.sph[3 * e + 1];
    sph[2 * P + e] = tb.sph[3 * e + 2];
  }
  double* cg = sph + 3 * P;
  for (int e = threadIdx.x; e < nT; e += blockDim.x) cg[e] = tb.CG[e];
  for (int e = threadIdx.x; e < LQRO_PRN * H; e += blockDim.x) cg[nT + e] = tb.prn[e];
  __syncthreads();
}

#endif  // LQRO_HULL_HOST_EMUL

// createHalfPlanes (MAIN:1085-1098) after `distance *= 0.5` (MAIN:1401); FP64 -> FP32 narrowing as in
// Vector3(double,double,double) (V3:113-118).
__device__ __forceinline__ void write_plane(float* pl, const double* vi, double distance, double n0, double n1, double n2, bool inside) {
  const double half = distance * 0.5;
  const double mult = inside ? 1.0 : -1.0;
  const double md = mult * half;
  pl[0] = (float)(vi[0] + md * n0);
  pl[1] = (float)(vi[1] + md * n1);
  pl[2] = (float)(vi[2] + md * n2);
  pl[3] = (float)n0;
  pl[4] = (float)n1;
  pl[5] = (float)n2;
}

// ------------------------------------------------------------------------------------------------
// Stage 2 of the pair kernel (k_pairs, lqro_step.cuh), one warp per pair: the pieces between "the surviving steps are known"
// and "GJK".  Plain device functions so that tests/pair_host can run the same source on the warp emulation.

// Translate_k of the surviving steps (klist[0..nk)), reference order (16-term left-to-right sums, MAT:218-231) -> tr[3k..3k+2].
// Terms whose state difference is exactly zero are skipped when that holds for all of entries 6..15: s + (+-0) == s bit for
// bit (s is never -0: the sum starts from +0), so the result is the reference's.
__device__ __forceinline__ void pair_translate_steps(const double* __restrict__ sNFt, const double* dx, const int* klist, int nk, int n3,
                                                     double* tr) {
  const int lane = lane_id();
  const bool tail0 = __ballot_sync(LQRO_FULL, lane >= 6 && lane < 16 && dx[lane] != 0.0) == 0u;
  for (int e = lane; e < 3 * nk; e += 32) {
    const int ki = e / 3, c = e - 3 * ki;
    const int k = klist[ki];
    const double* nf = sNFt + 3 * k + c;
    double s = 0;
#pragma unroll
    for (int m = 0; m < 6; ++m) s += nf[m * n3] * dx[m];
    if (!tail0) {
#pragma unroll
      for (int m = 6; m < 16; ++m) s += nf[m * n3] * dx[m];
    }
    tr[3 * k + c] = s;
  }
}

#ifndef LQRO_SCAN_STEPMAJOR
#define LQRO_SCAN_STEPMAJOR 1  // 0: the candidate-major walk also when P is a compile-time constant (the ablation the step-major walk was measured against)
#endif
// Candidates Transform_k (p_s + Translate_k) of the surviving steps + reachable filter (MAIN:687, 691-706) into the cloud
// (k-major, s ascending: the reference's order, MAIN:1383-1389); returns the number of reachable points.
// PT > 0: P known at compile time.
template <int PT>
__device__ __forceinline__ int pair_scan_cloud(const PairTables& t, const double* tr, const int* klist, int nk, int P, double vr0, double vr1,
                                               double vr2, double r2, double inv_r2, double r2lo, double r2hi, const CloudView& cl) {
  const int lane = lane_id();
  int count = 0;
  if constexpr (LQRO_SCAN_STEPMAJOR && PT > 0 && PT <= 64) {
    // Step-major walk.  A lane meets the same one or two samples p_s in every step (s = lane, lane + 32), so they stay in
    // registers for the whole pair, and Transform_k / Translate_k are read once per step instead of once per 32 candidates
    // (12 uniform loads per step against 16 per chunk).  Lanes past P in the last chunk of a step redo sample P-1 and are
    // masked out.  The body of a step is branch-free -- the chunks of a step are independent instruction streams the
    // scheduler can interleave (the kernel is bound by dependent-issue latency): candidates inside the decision band of
    // reach_test_band are only NOTED here; if there was one (practically never) the pair is redone by the walk below, whose
    // test evaluates the reference's own sequence in the band.
    constexpr int NCH = (PT + 31) / 32;
    double sx[NCH], sy[NCH], sz[NCH];
    bool live[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      live[c] = 32 * c + 32 <= PT || lane + 32 * c < PT;
      const int s = live[c] ? lane + 32 * c : PT - 1;
      sx[c] = t.sphx[s]; sy[c] = t.sphy[s]; sz[c] = t.sphz[s];
    }
    const unsigned lt = (1u << lane) - 1u;
    bool near = false;
    int k = nk > 0 ? klist[0] : 0;
    for (int ki = 0; ki < nk; ++ki) {
      const double* Tk = t.T + 9 * k;
      const double* trk = tr + 3 * k;
      k = klist[ki + 1 < nk ? ki + 1 : ki];  // the next step's index is on its way while this one is worked on
      const double T0 = Tk[0], T1 = Tk[1], T2 = Tk[2], T3 = Tk[3], T4 = Tk[4], T5 = Tk[5], T6 = Tk[6], T7 = Tk[7], T8 = Tk[8];
      const double tr0 = trk[0], tr1 = trk[1], tr2 = trk[2];
      double q0[NCH], q1[NCH], q2[NCH];
      unsigned mask[NCH];
      bool pass[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const double v0 = sx[c] + tr0, v1 = sy[c] + tr1, v2 = sz[c] + tr2;  // MAIN:687, the reference's operation order
        q0[c] = T0 * v0 + T1 * v1 + T2 * v2;
        q1[c] = T3 * v0 + T4 * v1 + T5 * v2;
        q2[c] = T6 * v0 + T7 * v1 + T8 * v2;
        const double a = q0[c] - vr0, b = q1[c] - vr1, cc = q2[c] - vr2;
        const double ss = fma(a, a, fma(b, b, cc * cc));
        pass[c] = ss < r2lo && live[c];
#ifdef LQRO_SCAN_TEST_WIDE_BAND  // emulation build only: note far more candidates than the band holds, so that the redo path runs
        near = near || (!(ss < 0.5 * r2lo) && !(ss > 2.0 * r2hi) && live[c]);
#else
        near = near || (!(ss < r2lo) && !(ss > r2hi) && live[c]);
#endif
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) mask[c] = __ballot_sync(LQRO_FULL, pass[c]);
      const bool fits = count + 32 * NCH <= cl.cap;  // warp-uniform: every point of this step lands in the shared part
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int pos = count + __popc(mask[c] & lt);
        if (pass[c]) {
          if (fits) { cl.sx[pos] = q0[c]; cl.sy[pos] = q1[c]; cl.sz[pos] = q2[c]; }
          else cl.put(pos, q0[c], q1[c], q2[c]);
        }
        count += __popc(mask[c]);
      }
    }
    if (!__any_sync(LQRO_FULL, near)) {
      __syncwarp();
      return count;
    }
    __syncwarp();
    count = 0;
  }
  {
    // candidate-major walk: 32 consecutive (step, sample) candidates per pass
    const int tot = nk * P;
    int ki = 0, s = lane;
    while (s >= P) { s -= P; ++ki; }
    for (int b0 = 0; b0 < tot; b0 += 32) {
      const int m = b0 + lane;
      // lanes past the end of the last chunk redo a valid candidate and are masked out: no divergent body
      double q0, q1, q2;
      candidate_point(t.T, t.sphx, t.sphy, t.sphz, tr, klist[ki < nk ? ki : nk - 1], s, q0, q1, q2);
      const bool pass = reach_test(q0 - vr0, q1 - vr1, q2 - vr2, r2, inv_r2) && m < tot;
      const unsigned mask = __ballot_sync(LQRO_FULL, pass);
      if (pass) {
        const int pos = count + __popc(mask & ((1u << lane) - 1u));
        if (count + 32 <= cl.cap) { cl.sx[pos] = q0; cl.sy[pos] = q1; cl.sz[pos] = q2; }  // warp-uniform fast path
        else cl.put(pos, q0, q1, q2);
      }
      count += __popc(mask);
      s += 32;
      { const bool wrap = s >= P; s -= wrap ? P : 0; ki += wrap ? 1 : 0; }
      if (P < 32) while (s >= P) { s -= P; ++ki; }
    }
  }
  __syncwarp();
  return count;
}

#ifndef LQRO_HULL_HOST_EMUL
// ------------------------------------------------------------------------------------------------
// Stage kernels behind the per-stage C entry points (one warp, one pair / one cloud).

// findFG (MAIN:653-662), ONE call: Atilde = A + (B*L), Btilde = B*E, F <- Atilde*F, G <- Atilde*G + Btilde, in place, every
// element the left-to-right sum Matrix::operator* forms (MAT:218-231).  256 threads, thread = element of the 16x16 product.
__global__ void __launch_bounds__(256) k_stage_findfg(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ L,
                                                      const double* __restrict__ E, double* __restrict__ F, double* __restrict__ G) {
  __shared__ double sAt[256], sF[256], sG[48];
  const int t = threadIdx.x, i = t >> 4, j = t & 15;
  {
    double bl = 0;
    for (int k = 0; k < 4; ++k) bl += B[i * 4 + k] * L[k * 16 + j];
    sAt[t] = A[t] + bl;
    sF[t] = F[t];
  }
  if (t < 48) sG[t] = G[t];
  __syncthreads();
  double f = 0;
  for (int m = 0; m < 16; ++m) f += sAt[i * 16 + m] * sF[m * 16 + j];
  F[t] = f;
  if (t < 48) {
    const int gi = t / 3, gj = t % 3;
    double be = 0, g = 0;
    for (int k = 0; k < 4; ++k) be += B[gi * 4 + k] * E[k * 3 + gj];
    for (int m = 0; m < 16; ++m) g += sAt[gi * 16 + m] * sG[m * 3 + gj];
    G[t] = g + be;
  }
}

// createObstacle (MAIN:683-688), ONE call with the caller's F, G and map P (3x16): Transform = !(P*G), Translate =
// -P*F*(xInit1-xInit2) (evaluated as ((-P)*F)*(xInit1-xInit2)), q_s = Transform*(points[s] + Translate).  One warp.
__global__ void __launch_bounds__(32) k_stage_create_obstacle(int n, const double* __restrict__ pts, const double* __restrict__ F,
                                                              const double* __restrict__ G, const double* __restrict__ P,
                                                              const double* __restrict__ x1, const double* __restrict__ x2,
                                                              double* __restrict__ Transform, double* __restrict__ Translate,
                                                              double* __restrict__ q) {
  __shared__ double sMPF[48], sT[9], sTr[3], sDx[16];
  const int lane = lane_id();
  if (lane < 16) sDx[lane] = x1[lane] - x2[lane];
  for (int e = lane; e < 48; e += 32) {   // (-P)*F, 3x16
    const int c = e >> 4, m = e & 15;
    double t = 0;
    for (int k = 0; k < 16; ++k) t += (-P[c * 16 + k]) * F[k * 16 + m];
    sMPF[e] = t;
  }
  if (lane == 0) {
    double pg[9];
    for (int e = 0; e < 9; ++e) {         // P*G, 3x3
      const int c = e / 3, r = e % 3;
      double t = 0;
      for (int k = 0; k < 16; ++k) t += P[c * 16 + k] * G[k * 3 + r];
      pg[e] = t;
    }
    inverse3_ref(pg, sT);
    for (int e = 0; e < 9; ++e) Transform[e] = sT[e];
  }
  __syncwarp();
  if (lane < 3) {
    double t = 0;
    for (int m = 0; m < 16; ++m) t += sMPF[lane * 16 + m] * sDx[m];
    sTr[lane] = t;
    Translate[lane] = t;
  }
  __syncwarp();
  for (int s = lane; s < n; s += 32) {
    const double v0 = pts[3 * s] + sTr[0], v1 = pts[3 * s + 1] + sTr[1], v2 = pts[3 * s + 2] + sTr[2];
    for (int r = 0; r < 3; ++r) {
      double t = 0;
      t += sT[3 * r] * v0; t += sT[3 * r + 1] * v1; t += sT[3 * r + 2] * v2;
      q[3 * s + r] = t;
    }
  }
}

// mode 0: all H*P candidate points of the pair (xi, xj) -> q_out (AoS [HP][3])
__global__ void __launch_bounds__(32) k_stage_candidates(LqroParams p, LqroTables tb, const double* __restrict__ xi,
                                                         const double* __restrict__ xj, double* __restrict__ q_out) {
  extern __shared__ double smem[];
  const int lane = lane_id();
  PairTables t;
  PairWarpSmem w;
  load_block_tables(smem, tb, p.H, p.P);
  carve_pair_smem(smem, 0, p.H, p.P, t, w);
  if (lane < 16) w.dx[lane] = xi[lane] - xj[lane];
  __syncwarp();
  warp_translate(t.NFt, w.dx, w.tr, p.H);
  __syncwarp();
  for (int m = lane; m < p.HP; m += 32) {
    const int k = m / p.P, s = m - k * p.P;
    double q0, q1, q2;
    candidate_point(t.T, t.sphx, t.sphy, t.sphz, w.tr, k, s, q0, q1, q2);
    q_out[3 * m] = q0; q_out[3 * m + 1] = q1; q_out[3 * m + 2] = q2;
  }
}

// reachable filter on an arbitrary cloud: idx_out ascending, *n_out count
__global__ void __launch_bounds__(32) k_stage_reachable(LqroParams p, const double* __restrict__ q, int n, const double* __restrict__ xi,
                                                        const double* __restrict__ xj, int* __restrict__ idx_out, int* __restrict__ n_out) {
  const int lane = lane_id();
  const double vr0 = xi[3] - xj[3], vr1 = xi[4] - xj[4], vr2 = xi[5] - xj[5];
  int count = 0;
  for (int base = 0; base < n; base += 32) {
    const int m = base + lane;
    bool pass = false;
    if (m < n) pass = reach_test(q[3 * m] - vr0, q[3 * m + 1] - vr1, q[3 * m + 2] - vr2, p.r2, p.inv_r2);
    const unsigned mask = __ballot_sync(LQRO_FULL, pass);
    if (pass) idx_out[count + __popc(mask & ((1u << lane) - 1u))] = m;
    count += __popc(mask);
  }
  if (lane == 0) *n_out = count;
}

// GJK on an arbitrary cloud (AoS in global memory, staged to SoA scratch)
__global__ void __launch_bounds__(32) k_stage_gjk(const double* __restrict__ vrel, const double* __restrict__ pts, int n,
                                                  double* __restrict__ scratch /*3n*/, double* __restrict__ out /*sqd, w2[3]*/,
                                                  int* __restrict__ iters) {
  const int lane = lane_id();
  double* px = scratch; double* py = scratch + n; double* pz = scratch + 2 * (size_t)n;
  for (int m = lane; m < n; m += 32) { px[m] = pts[3 * m]; py[m] = pts[3 * m + 1]; pz[m] = pts[3 * m + 2]; }
  __syncwarp();
  int it = 0;
  double w1[3], w2[3];
  CloudView cl;
  cl.sx = px; cl.sy = py; cl.sz = pz; cl.gx = cl.gy = cl.gz = nullptr; cl.cap = n;
#if LQRO_GJK_LANES
  __shared__ double gjk_scratch[LQRO_GJK_WARP_DOUBLES];
  const double sq = gjk_warp_lanes(vrel[0], vrel[1], vrel[2], cl, n, gjk_scratch, w1, w2, &it);
#else
  const double sq = gjk_warp(vrel[0], vrel[1], vrel[2], cl, n, w1, w2, &it);
#endif
  if (lane == 0) {
    out[0] = sq; out[1] = w2[0]; out[2] = w2[1]; out[3] = w2[2];
    out[4] = w1[0]; out[5] = w1[1]; out[6] = w1[2];
    *iters = it;
  }
}
#endif  // LQRO_HULL_HOST_EMUL
