// lqro_common.cuh -- shared definitions of the sm_100a LQR-Obstacles kernels.
//
// ARITHMETIC CONTRACT.  The whole library is compiled with -fmad=false: nvcc never contracts a*b+c,
// so every FP64/FP32 product and sum below is rounded separately, exactly like the reference built
// for x86-64 SSE2 (no FMA).  That is what makes the discrete decisions of the path (reachable masks,
// GJK simplex sequences, hull facet sets, LP branches) bit-identical to the CPU.  Double division and
// sqrt are IEEE-rounded on the device; float division/sqrt are too (-prec-div/-prec-sqrt default true,
// no --use_fast_math).  Where a kernel wants a fused multiply-add for speed on a path that provably
// cannot change a decision it calls fma()/fmaf() explicitly.
#pragma once
#ifndef LQRO_HULL_HOST_EMUL  // tests/*_host compile the kernel sources as plain C++ on a lockstep warp emulation
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define LQRO_XD 16
#define LQRO_WARP 32
#define LQRO_FULL 0xffffffffu

// Everything a kernel needs to know about the problem (filled by the host, passed by value).
struct LqroParams {
  int N;            // agents visible
  int local_begin;  // first owned agent
  int n_local;      // owned agents
  int K;            // neighbour slots per agent
  int H;            // horizon samples
  int P;            // ellipsoid samples
  int HP;           // H*P
  double max_speed; // reachable sphere radius
  double r2;        // max_speed^2 (the reference evaluates pow(xr,2) at run time: xr*xr)
  double inv_r2;    // 1/r2, only for the conservative fast test
  double r2lo, r2hi; // r2 (1 -+ 1e-9): decision band of reach_test_band (lqro_pair.cuh)
  float lp_radius;
  int hull_input;
  int fix_facet0;
  int first_slots;  // neighbour slots per agent in the first chunk of the slot-major pair enumeration (lqro_step.cuh)
  int prune;        // 1: skip horizon steps whose whole ellipsoid is provably unreachable (exact, see lqro_pair.cuh)
  double box[3];    // max_s |p_s| per coordinate: bounding box of the ellipsoid samples
  int seg_split;    // > 0: the owned agents are enumerated as two segments, [0, seg_split) first (lqro_enum.cuh: neighbour rows still arriving)
};

// Device tables of one dynamics class (pair independent, built once by k_horizon_tables):
//   NFt  [16][3H]  NFt[m][k*3+c] = -(C F_k)(c, m)       (transposed for conflict-free shared loads)
//   T    [H][9]    (C G_k)^-1, row major
//   sph  [P][3]    samples of the doubled ellipsoid
//   CG   [H][9]    C G_k (the matrix T_k inverts), row major
//   prn  [H][8]    pruning constants of step k (see step_prunable): sigma_max(C G_k) (upper bound; stored as (max_speed*sigma_max)/max_speed),
//                  kappa_k = |T_k|_F*sigma_max(C G_k), |T_k(0,0)|, |T_k(1,1)|, |T_k(2,2)|, |offdiag(T_k)|_F, |T_k|_F, 0
struct LqroTables {
  const double* NFt;
  const double* T;
  const double* sph;
  const double* CG;
  const double* prn;
};

#ifndef LQRO_HULL_HOST_EMUL
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

// 3-term dot product in the reference's association: (a0*b0 + a1*b1) + a2*b2   (GJK.h:36-38, MAT:218-231)
__device__ __forceinline__ double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
  return a0 * b0 + a1 * b1 + a2 * b2;
}
