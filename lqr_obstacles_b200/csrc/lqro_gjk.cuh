// lqro_gjk.cuh -- point-versus-cloud GJK distance, replaying Cameron's v2.4 routine exactly as the
// reference calls it (run_gjk MAIN:712-746 -> gjk_distance GJK.c:296-501): object 1 is the single
// point v_rel, object 2 the reachable cloud, identity transforms, no seed, both witnesses.
//
// The routine is not an exact distance solver (it stops on the absolute test g_val < 1e-8 on a
// SQUARED distance, GJK.c:465), so parity to 1e-9 requires replaying its arithmetic, not solving
// the geometry: same subset order, same delta recursion, same operation order, no FMA.
//
// Mapping: one warp per pair.  The support scan over the cloud (the only O(n) part) is spread over
// the 32 lanes and reduced with shuffles to (max value, lowest index) -- the reference's strict '>'
// scan (GJK.c:786) keeps the lowest index on ties.  The simplex sub-algorithm (Johnson's deltas over
// <= 15 subsets) is scalar work: lane 0 runs it on a per-warp state block in shared memory.
#pragma once
#include "lqro_common.cuh"

// Subsets of the <=4 simplex points as bit masks (element lists ascending): cardinality, elements
// and non-elements.  predecessor(s, j) = s & ~(1 << elt(s, j)), successor(s, k) = s | (1 << nonelt(s, k)).
__constant__ int c_gjk_card[16] = {0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
__constant__ signed char c_gjk_elt[16][4] = {
    {0, 0, 0, 0}, {0, 0, 0, 0}, {1, 0, 0, 0}, {0, 1, 0, 0}, {2, 0, 0, 0}, {0, 2, 0, 0}, {1, 2, 0, 0}, {0, 1, 2, 0},
    {3, 0, 0, 0}, {0, 3, 0, 0}, {1, 3, 0, 0}, {0, 1, 3, 0}, {2, 3, 0, 0}, {0, 2, 3, 0}, {1, 2, 3, 0}, {0, 1, 2, 3}};
__constant__ signed char c_gjk_nonelt[16][4] = {
    {0, 1, 2, 3}, {1, 2, 3, 0}, {0, 2, 3, 0}, {2, 3, 0, 0}, {0, 1, 3, 0}, {1, 3, 0, 0}, {0, 3, 0, 0}, {3, 0, 0, 0},
    {0, 1, 2, 0}, {1, 2, 0, 0}, {0, 2, 0, 0}, {2, 0, 0, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};

#define LQRO_GJK_EPS 1.0e-8    // GJK.h:208
#define LQRO_GJK_TINY 1.0e-20  // GJK.h:214

struct GjkState {
  double c2[4][3];     // cloud-side coordinates of the simplex points (the other side is always v_rel)
  double lambdas[4];
  double del[16][4];   // delta_values[s][element]
  double dsum[16];     // delta[s]
  double prod[4][4];
  double w1[3], w2[3]; // witness points: on {v_rel}, on conv(cloud)
  double disp[3];
  double sqrd;
  int npts;
  int use_default;
  int simplex2[4];
};

// GJK.c:527-581
__device__ __forceinline__ void gjk_compute_subterms(GjkState& g, const double* vrel) {
  const int size = g.npts;
  double c[4][3];
  for (int i = 0; i < size; ++i)
    for (int j = 0; j < 3; ++j) c[i][j] = vrel[j] - g.c2[i][j];
  for (int i = 0; i < size; ++i)
    for (int j = i; j < size; ++j) {
      double v = dot3(c[i][0], c[i][1], c[i][2], c[j][0], c[j][1], c[j][2]);
      g.prod[i][j] = v;
      g.prod[j][i] = v;
    }
  const int smax = 1 << size;
  for (int s = 1; s < smax; ++s) {
    const int cs = c_gjk_card[s];
    if (cs <= 1) { g.del[s][c_gjk_elt[s][0]] = 1.0; continue; }
    if (cs == 2) {
      const int e0 = c_gjk_elt[s][0], e1 = c_gjk_elt[s][1];
      g.del[s][e0] = g.prod[e1][e1] - g.prod[e1][e0];
      g.del[s][e1] = g.prod[e0][e0] - g.prod[e0][e1];
      continue;
    }
    for (int j = 0; j < cs; ++j) {
      const int jelt = c_gjk_elt[s][j];
      const int jsub = s & ~(1 << jelt);
      const int e0 = c_gjk_elt[jsub][0];
      const int cj = c_gjk_card[jsub];
      double sum = 0;
      for (int i = 0; i < cj; ++i) {
        const int ielt = c_gjk_elt[jsub][i];
        sum += g.del[jsub][ielt] * (g.prod[ielt][e0] - g.prod[ielt][jelt]);
      }
      g.del[s][jelt] = sum;
    }
  }
}

// GJK.c:708-736
__device__ __forceinline__ void gjk_reset_simplex(GjkState& g, int subset) {
  const int cs = c_gjk_card[subset];
  for (int j = 0; j < cs; ++j) {
    const int oldpos = c_gjk_elt[subset][j];
    if (oldpos != j) {
      g.simplex2[j] = g.simplex2[oldpos];
      for (int i = 0; i < 3; ++i) g.c2[j][i] = g.c2[oldpos][i];
    }
    g.lambdas[j] = g.del[subset][oldpos] / g.dsum[subset];
  }
  g.npts = cs;
}

// GJK.c:593-657
__device__ __forceinline__ int gjk_default_distance(GjkState& g) {
  const int size = g.npts;
  const int smax = 1 << size;
  int s, ok = 0;
  for (s = 1; s < smax; ++s) {
    g.dsum[s] = 0;
    ok = 1;
    const int cs = c_gjk_card[s];
    for (int j = 0; ok && j < cs; ++j) {
      const double d = g.del[s][c_gjk_elt[s][j]];
      if (d > 0) g.dsum[s] += d; else ok = 0;
    }
    for (int k = 0; ok && k < size - cs; ++k) {
      const int ne = c_gjk_nonelt[s][k];
      if (g.del[s | (1 << ne)][ne] > 0) ok = 0;
    }
    if (ok && g.dsum[s] >= LQRO_GJK_TINY) break;
  }
  if (ok && s < smax) { gjk_reset_simplex(g, s); return 1; }
  return 0;
}

// GJK.c:663-706
__device__ __forceinline__ void gjk_backup_distance(GjkState& g) {
  const int size = g.npts;
  const int smax = 1 << size;
  int bests = 0;
  double num[16], den[16];
  for (int s = 1; s < smax; ++s) {
    if (g.dsum[s] <= 0) continue;
    const int cs = c_gjk_card[s];
    int i;
    for (i = 0; i < cs; ++i)
      if (g.del[s][c_gjk_elt[s][i]] <= 0) break;
    if (i < cs) continue;
    double acc = 0;
    for (int j = 0; j < cs; ++j)
      for (int k = 0; k < cs; ++k) {
        const int ej = c_gjk_elt[s][j], ek = c_gjk_elt[s][k];
        acc += (g.del[s][ej] * g.del[s][ek]) * g.prod[ej][ek];
      }
    num[s] = acc;
    den[s] = g.dsum[s] * g.dsum[s];
    if (bests < 1 || num[s] * den[bests] < num[bests] * den[s]) bests = s;
  }
  gjk_reset_simplex(g, bests);
}

// Scalar half of one main-loop iteration (GJK.c:374-413): sub-algorithm, witnesses, displacement.
__device__ __forceinline__ void gjk_scalar_step(GjkState& g, const double* vrel) {
  if (g.npts == 1) {
    g.lambdas[0] = 1.0;
  } else {
    gjk_compute_subterms(g, vrel);
    if (g.use_default) g.use_default = gjk_default_distance(g);
    if (!g.use_default) gjk_backup_distance(g);
  }
  for (int d = 0; d < 3; ++d) {  // compute_point, GJK.c:851-862 (accumulated from 0)
    double a = 0, b = 0;
    for (int i = 0; i < g.npts; ++i) a += vrel[d] * g.lambdas[i];
    for (int i = 0; i < g.npts; ++i) b += g.c2[i][d] * g.lambdas[i];
    g.w1[d] = a;
    g.w2[d] = b;
  }
  for (int d = 0; d < 3; ++d) g.disp[d] = g.w2[d] - g.w1[d];
  g.sqrd = dot3(g.disp[0], g.disp[1], g.disp[2], g.disp[0], g.disp[1], g.disp[2]);
}

// Warp-cooperative GJK.  px/py/pz: the cloud (shared or global), n points.  Returns d^2 (uniform
// across the warp); witnesses are left in g.w1 / g.w2; *iters receives the number of G-tests.
__device__ __forceinline__ double gjk_warp(GjkState& g, const double vrel0, const double vrel1, const double vrel2,
                                           const double* __restrict__ px, const double* __restrict__ py,
                                           const double* __restrict__ pz, int n, int* iters) {
  const int lane = lane_id();
  const double vrel[3] = {vrel0, vrel1, vrel2};
  if (lane == 0) {  // GJK.c:340-347
    g.npts = 1;
    g.simplex2[0] = 0;
    g.lambdas[0] = 1.0;
    g.c2[0][0] = px[0]; g.c2[0][1] = py[0]; g.c2[0][2] = pz[0];
    g.use_default = 1;
  }
  __syncwarp();
  int first_iteration = 1;
  double oldsqrd = 0.0, sqrd = 0.0;
  int max_iterations = n;  // NumVertices(obj1) * NumVertices(obj2), GJK.c:367
  int ntests = 0;
  bool exhausted = true;
  while (max_iterations-- > 0) {
    if (lane == 0) gjk_scalar_step(g, vrel);
    __syncwarp();
    sqrd = g.sqrd;
    if (sqrd < LQRO_GJK_EPS) { exhausted = false; break; }  // GJK.c:419-422
    const double d0 = g.disp[0], d1 = g.disp[1], d2 = g.disp[2];
    const double r0 = -d0, r1 = -d1, r2 = -d2;  // reverse_displacementv
    // support of the cloud in direction r (support_simple, GJK.c:770-794)
    double best = -1.7976931348623157e308;
    int besti = 0x7fffffff;
    for (int m = lane; m < n; m += 32) {
      const double v = dot3(px[m], py[m], pz[m], r0, r1, r2);
      if (v > best) { best = v; besti = m; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(LQRO_FULL, best, off);
      const int oi = __shfl_xor_sync(LQRO_FULL, besti, off);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    const double minus_minv = best;
    const int minp = besti;
    const double maxv = dot3(vrel0, vrel1, vrel2, d0, d1, d2);  // support of the single point
    ++ntests;
    double g_val = sqrd + maxv + minus_minv;  // GJK.c:450
    if (g_val < 0.0) g_val = 0;
    if (g_val < LQRO_GJK_EPS) { exhausted = false; break; }
    const int npts = g.npts;
    if ((first_iteration || sqrd < oldsqrd) && npts <= 3) {  // GJK.c:472-487
      __syncwarp();
      if (lane == 0) {
        g.simplex2[npts] = minp;
        g.lambdas[npts] = 0.0;
        g.c2[npts][0] = px[minp]; g.c2[npts][1] = py[minp]; g.c2[npts][2] = pz[minp];
        g.npts = npts + 1;
        g.use_default = 1;
      }
      __syncwarp();
      oldsqrd = sqrd;
      first_iteration = 0;
      continue;
    }
    const int use_default = g.use_default;  // GJK.c:490-496
    __syncwarp();
    if (use_default) {
      if (lane == 0) g.use_default = 0;
      __syncwarp();
    } else {
      exhausted = false;
      break;
    }
  }
  if (iters) *iters = ntests;
  if (exhausted) return 0.0;  // GJK.c:499
  return sqrd;
}
