// lqro_gjk.cuh -- point-versus-cloud GJK distance, replaying Cameron's v2.4 routine exactly as the
// reference calls it (run_gjk MAIN:712-746 -> gjk_distance GJK.c:296-501): object 1 is the single
// point v_rel, object 2 the reachable cloud, identity transforms, no seed, both witnesses.
//
// The routine is not an exact distance solver (it stops on the absolute test g_val < 1e-8 on a
// SQUARED distance, GJK.c:465), so parity to 1e-9 requires replaying its arithmetic, not solving
// the geometry: same subset order, same delta recursion, same operation order, no FMA.
//
// Mapping: one warp per pair.  The support scan over the cloud (the only O(n) part) is spread over
// the 32 lanes and reduced to (max value, lowest index) -- the reference's strict '>' scan (GJK.c:786)
// keeps the lowest index on ties -- with three REDUX operations on order-preserving keys.  The simplex
// sub-algorithm (Johnson's deltas over <= 15 subsets) runs LANE-PARALLEL by default (gjk_warp_lanes: one
// (subset, element) pair per lane, see below).  The older scalar version (gjk_warp: every lane runs it
// redundantly on register-resident state, subset loops templated on the simplex size and fully unrolled)
// is kept behind LQRO_GJK_LANES=0 as the ablation it was measured against; both are bit-identical.
#pragma once
#include "lqro_common.cuh"

#ifndef LQRO_GJK_LANES
#define LQRO_GJK_LANES 1       // 1: lane-parallel Johnson sub-algorithm (default); 0: register-resident scalar version
#endif
#ifndef LQRO_GJK_STEP2
#define LQRO_GJK_STEP2 1       // 1: the two-point default step written out (gjkl_default_step2); 0: the general lane-parallel step for every simplex size
#endif
#ifndef LQRO_GJK_STEP3
#define LQRO_GJK_STEP3 0       // 1: the three-point default step on the register-resident templates (gjkl_default_step_regs<3>); measured slower (+0.10 ms)
#endif
#ifndef LQRO_GJK_STEP4
#define LQRO_GJK_STEP4 0       // 1: the four-point default step too
#endif
#define LQRO_GJK_EPS 1.0e-8    // GJK.h:208
#define LQRO_GJK_TINY 1.0e-20  // GJK.h:214

// Subsets of the <= 4 simplex points as bit masks; element lists ascending (the tables of GJK.c:88-159).
// predecessor(s, j) = s & ~(1 << elt(s, j)), successor(s, k) = s | (1 << nonelt(s, k)).
__host__ __device__ constexpr int gjk_card(int s) { return (s & 1) + ((s >> 1) & 1) + ((s >> 2) & 1) + ((s >> 3) & 1); }
__host__ __device__ constexpr int gjk_elt(int s, int i) {
  int c = 0;
  for (int b = 0; b < 4; ++b)
    if (s & (1 << b)) { if (c == i) return b; ++c; }
  return 0;
}
__host__ __device__ constexpr int gjk_nonelt(int s, int i) {
  int c = 0;
  for (int b = 0; b < 4; ++b)
    if (!(s & (1 << b))) { if (c == i) return b; ++c; }
  return 0;
}

struct GjkRegs {
  double c2[4][3];   // cloud-side coordinates of the simplex points (the other side is always v_rel)
  double lam[4];
  double del[16][4]; // delta_values[s][element]
  double dsum[16];   // delta[s]
  double prod[4][4];
  double w1[3], w2[3];
  int npts;
  int use_default;
};

// GJK.c:527-581
template <int SIZE>
__device__ __forceinline__ void gjk_subterms(GjkRegs& g, double v0, double v1, double v2) {
  double c[SIZE][3];
#pragma unroll
  for (int i = 0; i < SIZE; ++i) { c[i][0] = v0 - g.c2[i][0]; c[i][1] = v1 - g.c2[i][1]; c[i][2] = v2 - g.c2[i][2]; }
#pragma unroll
  for (int i = 0; i < SIZE; ++i)
#pragma unroll
    for (int j = i; j < SIZE; ++j) {
      const double v = dot3(c[i][0], c[i][1], c[i][2], c[j][0], c[j][1], c[j][2]);
      g.prod[i][j] = v;
      g.prod[j][i] = v;
    }
#pragma unroll
  for (int s = 1; s < (1 << SIZE); ++s) {
    const int cs = gjk_card(s);
    if (cs <= 1) {
      g.del[s][gjk_elt(s, 0)] = 1.0;
    } else if (cs == 2) {
      const int e0 = gjk_elt(s, 0), e1 = gjk_elt(s, 1);
      g.del[s][e0] = g.prod[e1][e1] - g.prod[e1][e0];
      g.del[s][e1] = g.prod[e0][e0] - g.prod[e0][e1];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < cs) {
          const int jelt = gjk_elt(s, j);
          const int jsub = s & ~(1 << jelt);
          const int e0 = gjk_elt(jsub, 0);
          double sum = 0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            if (i < gjk_card(jsub)) {
              const int ielt = gjk_elt(jsub, i);
              sum += g.del[jsub][ielt] * (g.prod[ielt][e0] - g.prod[ielt][jelt]);
            }
          }
          g.del[s][jelt] = sum;
        }
      }
    }
  }
}

// GJK.c:708-736 for a compile-time subset
template <int S>
__device__ __forceinline__ void gjk_reset_to(GjkRegs& g) {
  constexpr int cs = gjk_card(S);
#pragma unroll
  for (int j = 0; j < cs; ++j) {
    const int oldpos = gjk_elt(S, j);
    if (oldpos != j) { g.c2[j][0] = g.c2[oldpos][0]; g.c2[j][1] = g.c2[oldpos][1]; g.c2[j][2] = g.c2[oldpos][2]; }
    g.lam[j] = g.del[S][oldpos] / g.dsum[S];
  }
  g.npts = cs;
}
__device__ __forceinline__ void gjk_reset_simplex(GjkRegs& g, int subset) {
  switch (subset) {
    case 1: gjk_reset_to<1>(g); break;   case 2: gjk_reset_to<2>(g); break;   case 3: gjk_reset_to<3>(g); break;
    case 4: gjk_reset_to<4>(g); break;   case 5: gjk_reset_to<5>(g); break;   case 6: gjk_reset_to<6>(g); break;
    case 7: gjk_reset_to<7>(g); break;   case 8: gjk_reset_to<8>(g); break;   case 9: gjk_reset_to<9>(g); break;
    case 10: gjk_reset_to<10>(g); break; case 11: gjk_reset_to<11>(g); break; case 12: gjk_reset_to<12>(g); break;
    case 13: gjk_reset_to<13>(g); break; case 14: gjk_reset_to<14>(g); break; case 15: gjk_reset_to<15>(g); break;
    default: break;
  }
}

// GJK.c:593-657 without the reset: the first viable subset in numeric order (0: none); delta[s] of every subset the scan reaches
template <int SIZE>
__device__ __forceinline__ int gjk_default_select(GjkRegs& g) {
  int sel = 0;
#pragma unroll
  for (int s = 1; s < (1 << SIZE); ++s) {
    if (sel == 0) {
      const int cs = gjk_card(s);
      double ds = 0;
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < cs && ok) {
          const double d = g.del[s][gjk_elt(s, j)];
          if (d > 0) ds += d; else ok = false;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < SIZE - cs && ok) {
          const int ne = gjk_nonelt(s, k);
          if (g.del[s | (1 << ne)][ne] > 0) ok = false;
        }
      }
      g.dsum[s] = ds;
      if (ok && ds >= LQRO_GJK_TINY) sel = s;
    }
  }
  return sel;
}
// GJK.c:593-657.  Returns 1 when a viable subset was found (and the simplex reset to it).
template <int SIZE>
__device__ __forceinline__ int gjk_default_distance(GjkRegs& g) {
  const int sel = gjk_default_select<SIZE>(g);
  if (sel) { gjk_reset_simplex(g, sel); return 1; }
  return 0;
}

// GJK.c:663-706
template <int SIZE>
__device__ __forceinline__ void gjk_backup_distance(GjkRegs& g) {
  int bests = 0;
  double bnum = 0, bden = 0;
#pragma unroll
  for (int s = 1; s < (1 << SIZE); ++s) {
    const int cs = gjk_card(s);
    bool viable = g.dsum[s] > 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < cs && g.del[s][gjk_elt(s, i)] <= 0) viable = false;
    if (viable) {
      double acc = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (j < cs && k < cs) {
            const int ej = gjk_elt(s, j), ek = gjk_elt(s, k);
            acc += (g.del[s][ej] * g.del[s][ek]) * g.prod[ej][ek];
          }
      const double den = g.dsum[s] * g.dsum[s];
      if (bests < 1 || acc * bden < bnum * den) { bests = s; bnum = acc; bden = den; }
    }
  }
  gjk_reset_simplex(g, bests);
}

template <int SIZE>
__device__ __forceinline__ void gjk_simplex_step(GjkRegs& g, double v0, double v1, double v2) {
  gjk_subterms<SIZE>(g, v0, v1, v2);
  if (g.use_default) g.use_default = gjk_default_distance<SIZE>(g);
  if (!g.use_default) gjk_backup_distance<SIZE>(g);
}

template <int N>
__device__ __forceinline__ void gjk_witness(GjkRegs& g, double v0, double v1, double v2) {
  const double vv[3] = {v0, v1, v2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) a += vv[d] * g.lam[i];
#pragma unroll
    for (int i = 0; i < N; ++i) b += g.c2[i][d] * g.lam[i];
    g.w1[d] = a;
    g.w2[d] = b;
  }
}

// The cloud of one pair: the first `cap` points in shared memory (SoA), the rest in a per-warp overflow
// area in global memory (rare: dense scenes).  For a plain global cloud set cap = n.
struct CloudView {
  double *sx, *sy, *sz;   // [cap]
  double *gx, *gy, *gz;   // [n - cap], may be null when n <= cap
  int cap;
  __device__ __forceinline__ void get(int m, double& x, double& y, double& z) const {
    if (m < cap) { x = sx[m]; y = sy[m]; z = sz[m]; }
    else { x = gx[m - cap]; y = gy[m - cap]; z = gz[m - cap]; }
  }
  __device__ __forceinline__ void put(int m, double x, double y, double z) const {
    if (m < cap) { sx[m] = x; sy[m] = y; sz[m] = z; }
    else { gx[m - cap] = x; gy[m - cap] = y; gz[m - cap] = z; }
  }
};

// ------------------------------------------------------------------------------------------------
// LANE-PARALLEL Johnson sub-algorithm (the default).  A simplex of <= 4 points has 15 non-empty subsets and
// sum |s| = 32 (subset, element) pairs: exactly one per lane.  Lane (s, j) owns delta_values[s][j]; the
// recursion over subset size (GJK.c:527-581) proceeds level by level with shuffles from the lanes that own the
// predecessors' deltas, the dot products live in shared memory (16 doubles per warp, next to the 4 simplex
// points), the sign tests of GJK.c:593-657 become one ballot plus mask arithmetic, and "first viable subset in
// numeric order" is an OR-reduction.  Every delta is produced by the same operations in the same order as in the
// scalar routine, so results are bit-identical to it; what changes is that no lane carries the 15-subset
// tables in registers (the register-resident version needs ~160 registers at simplex size 4 and spills under
// the 128-register budget of the 16-warp pair kernel).
//   lane = e*8 + idx, idx = s with bit e squeezed out  <=>  (s, e) with bit e of s set.
__device__ __forceinline__ int gjk_lane_of(int s, int e) { return e * 8 + ((s & ((1 << e) - 1)) | ((s >> (e + 1)) << e)); }
#define LQRO_GJK_WARP_DOUBLES 28  // sc2[4][3] + sprod[4][4]

struct GjkLane {
  double* sc2;    // [4][3] cloud-side coordinates of the simplex points (the other side is always v_rel)
  double* sprod;  // [4][4] dot products (GJK.c:540-548)
  double lam[4];  // uniform
  double del;     // delta_values[my_s][my_j]
  double dsum;    // delta[my_s], kept by the lane of the subset's lowest element; persists across iterations like
                  // the reference's static tables (the backup procedure may read values of an earlier iteration)
  double de[4];   // deltas of my own subset, by element (gathered after the recursion)
  int my_s, my_j;
  unsigned sup[3];  // sup[size-2]: lanes (s + k, k), k < size not in my_s -- a positive delta there rules my subset out (GJK.c:624-631)
  int npts, use_default;
};

// GJK.c:527-581 for a simplex of `size` points
__device__ __forceinline__ void gjkl_subterms(GjkLane& g, int size, int lane, double v0, double v1, double v2) {
  if (lane < 16) {
    const int a = lane >> 2, b = lane & 3;
    if (a < size && b < size) {
      const int i = a < b ? a : b, j = a < b ? b : a;
      const double ci0 = v0 - g.sc2[3 * i], ci1 = v1 - g.sc2[3 * i + 1], ci2 = v2 - g.sc2[3 * i + 2];
      const double cj0 = v0 - g.sc2[3 * j], cj1 = v1 - g.sc2[3 * j + 1], cj2 = v2 - g.sc2[3 * j + 2];
      g.sprod[lane] = dot3(ci0, ci1, ci2, cj0, cj1, cj2);
    }
  }
  __syncwarp();
  const int s = g.my_s, j = g.my_j;
  const int cs = __popc(s);
  const bool act = (s >> size) == 0;
  double del = 1.0;
  if (act && cs == 2) {
    const int k = __ffs(s & ~(1 << j)) - 1;
    del = g.sprod[k * 4 + k] - g.sprod[k * 4 + j];
  }
  const int jsub = s & ~(1 << j);
  const int e0 = jsub ? __ffs(jsub) - 1 : 0;
  for (int c = 3; c <= size; ++c) {
    const bool mine = act && cs == c;
    double sum = 0;
    int rem = jsub;
    for (int t = 0; t < c - 1; ++t) {
      const int i = mine ? __ffs(rem) - 1 : 0;
      const int src = mine ? gjk_lane_of(jsub, i) : lane;
      const double dv = __shfl_sync(LQRO_FULL, del, src);
      if (mine) { sum += dv * (g.sprod[i * 4 + e0] - g.sprod[i * 4 + j]); rem &= rem - 1; }
    }
    if (mine) del = sum;
  }
  g.del = del;
  // every lane collects the deltas of its own subset, by element
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const bool in = (s >> e) & 1;
    const double dv = __shfl_sync(LQRO_FULL, del, in ? gjk_lane_of(s, e) : lane);
    g.de[e] = dv;
  }
}

// GJK.c:708-736: the simplex becomes subset `sel` (elements move down in order), lambdas = delta_values / delta
__device__ __forceinline__ void gjkl_reset(GjkLane& g, int sel, int lane) {
  const double ds = __shfl_sync(LQRO_FULL, g.dsum, gjk_lane_of(sel, __ffs(sel) - 1));
  int rem = sel;
  const int cs = __popc(sel);
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int oldpos = jj < cs ? __ffs(rem) - 1 : 0;
    const double dv = __shfl_sync(LQRO_FULL, g.del, jj < cs ? gjk_lane_of(sel, oldpos) : lane);
    if (jj < cs) {
      rem &= rem - 1;
      if (lane < 3 && oldpos != jj) g.sc2[3 * jj + lane] = g.sc2[3 * oldpos + lane];  // sources lie at or above jj: in-place is safe
      g.lam[jj] = dv / ds;
    }
  }
  g.npts = cs;
  __syncwarp();
}

// GJK.c:593-657.  Returns 1 when a viable subset was found (and the simplex reset to it).
__device__ __forceinline__ int gjkl_default_distance(GjkLane& g, int size, int lane) {
  const int s = g.my_s;
  const bool act = (s >> size) == 0;
  const unsigned P = __ballot_sync(LQRO_FULL, act && g.del > 0);
  bool ok = true;
  double ds = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (((s >> e) & 1) && ok) { if (g.de[e] > 0) ds += g.de[e]; else ok = false; }
  if (P & (size == 2 ? g.sup[0] : size == 3 ? g.sup[1] : g.sup[2])) ok = false;
  const bool valid = act && ok && ds >= LQRO_GJK_TINY;
  const unsigned V = __reduce_or_sync(LQRO_FULL, valid ? 1u << s : 0u);
  const int sel = V ? __ffs(V) - 1 : 0;
  // the scalar loop assigns delta[s] for every subset it reaches: all of them, or up to the selected one
  if (act && g.my_j == __ffs(s) - 1 && (sel == 0 || s <= sel)) g.dsum = ds;
  if (sel) { gjkl_reset(g, sel, lane); return 1; }
  return 0;
}

// GJK.c:663-706
__device__ __forceinline__ void gjkl_backup_distance(GjkLane& g, int size, int lane) {
  const int s = g.my_s;
  const bool act = (s >> size) == 0;
  const int low = __ffs(s) - 1;
  const double mydsum = __shfl_sync(LQRO_FULL, g.dsum, gjk_lane_of(s, low));
  bool viable = act && mydsum > 0;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (((s >> e) & 1) && g.de[e] <= 0) viable = false;
  double acc = 0;
#pragma unroll
  for (int ej = 0; ej < 4; ++ej)
#pragma unroll
    for (int ek = 0; ek < 4; ++ek)
      if (act && ((s >> ej) & 1) && ((s >> ek) & 1)) acc += (g.de[ej] * g.de[ek]) * g.sprod[ej * 4 + ek];
  const double den = mydsum * mydsum;
  int bests = 0;
  double bnum = 0, bden = 0;
  for (int t = 1; t < (1 << size); ++t) {  // the reference's order-dependent scan, uniform on all lanes
    const int src = gjk_lane_of(t, __ffs(t) - 1);
    const int tv = __shfl_sync(LQRO_FULL, (int)viable, src);
    const double ta = __shfl_sync(LQRO_FULL, acc, src);
    const double td = __shfl_sync(LQRO_FULL, den, src);
    if (tv && (bests < 1 || ta * bden < bnum * td)) { bests = t; bnum = ta; bden = td; }
  }
  if (bests) gjkl_reset(g, bests, lane);
}

template <int N>
__device__ __forceinline__ void gjkl_witness(const GjkLane& g, double v0, double v1, double v2, double* w1, double* w2) {
  const double vv[3] = {v0, v1, v2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) a += vv[d] * g.lam[i];
#pragma unroll
    for (int i = 0; i < N; ++i) b += g.sc2[3 * i + d] * g.lam[i];
    w1[d] = a;
    w2[d] = b;
  }
}

__device__ __forceinline__ void gjkl_simplex_step(GjkLane& g, int lane, double v0, double v1, double v2) {
  const int size = g.npts;
  gjkl_subterms(g, size, lane, v0, v1, v2);
  if (g.use_default) g.use_default = gjkl_default_distance(g, size, lane);
  if (!g.use_default) gjkl_backup_distance(g, size, lane);
}

// The default procedure for a simplex of TWO points (every GJK run passes through it once), written out: 3 subsets,
// 2 deltas that are not 1 -- cheaper for every lane to evaluate redundantly than the lane-parallel machinery above.
// Same operations in the same order as gjk_subterms<2> + gjk_default_distance<2> + gjk_reset_to (GJK.c:527-657, 708-736):
//   delta_3[0] = prod[1][1] - prod[1][0], delta_3[1] = prod[0][0] - prod[0][1];
//   subset {0} is chosen when delta_3[1] <= 0, else {1} when delta_3[0] <= 0, else {0,1} when the sum reaches TINY.
// Returns false WITHOUT touching any state when no subset is viable (the caller then runs the general step, whose backup
// procedure needs the full tables); otherwise the simplex is reset and delta[s] of the subsets the scan reached is recorded in
// the owning lanes, as the general step does (the backup procedure of a later iteration may read them).
__device__ __forceinline__ bool gjkl_default_step2(GjkLane& g, int lane, double v0, double v1, double v2) {
  const double a0 = v0 - g.sc2[0], a1 = v1 - g.sc2[1], a2 = v2 - g.sc2[2];
  const double b0 = v0 - g.sc2[3], b1 = v1 - g.sc2[4], b2 = v2 - g.sc2[5];
  const double p00 = dot3(a0, a1, a2, a0, a1, a2), p01 = dot3(a0, a1, a2, b0, b1, b2), p11 = dot3(b0, b1, b2, b0, b1, b2);
  const double d30 = p11 - p01, d31 = p00 - p01;
  if (!(d31 > 0)) {                 // subset 1 = {0}: its own delta is 1, the successor's delta of point 1 is not positive
    if (lane == 0) g.dsum = 1.0;    // lane_of(1, 0)
    g.lam[0] = 1.0;                 // 1 / 1
    g.npts = 1;
    return true;
  }
  if (!(d30 > 0)) {                 // subset 2 = {1}
    if (lane == 0 || lane == 8) g.dsum = 1.0;  // lane_of(1, 0), lane_of(2, 1)
    __syncwarp();
    if (lane < 3) g.sc2[lane] = g.sc2[3 + lane];
    g.lam[0] = 1.0;
    g.npts = 1;
    __syncwarp();
    return true;
  }
  const double ds = d30 + d31;      // (0 + d30) + d31, both positive
  if (!(ds >= LQRO_GJK_TINY)) return false;
  if (lane == 0 || lane == 8) g.dsum = 1.0;
  if (lane == 1) g.dsum = ds;       // lane_of(3, 0)
  g.lam[0] = d30 / ds;
  g.lam[1] = d31 / ds;
  g.npts = 2;
  return true;
}

// The default procedure for a simplex of SIZE = 3 (or 4) points evaluated by every lane redundantly with the register-resident
// templates of the scalar routine (gjk_subterms / gjk_default_select: subset loops unrolled at compile time, the same operations
// in the same order), fed from and written back to the lane-parallel state.  The lane-parallel recursion pays a chain of
// dependent shuffles per subset size (and its collectives per test); this pays FP64 instructions (63 for three points) with no
// dependence between most of them -- the pair kernel is bound by dependent-issue latency, so this is the cheaper currency.
// Returns false WITHOUT touching any state when no subset is viable (the caller runs the general step and its backup procedure).
template <int S>
__device__ __forceinline__ void gjk_pick_subset(const GjkRegs& r, double& ds, double* num) {
  ds = r.dsum[S];
#pragma unroll
  for (int j = 0; j < gjk_card(S); ++j) num[j] = r.del[S][gjk_elt(S, j)];
}
template <int SIZE>
__device__ __forceinline__ bool gjkl_default_step_regs(GjkLane& g, int lane, double v0, double v1, double v2) {
  GjkRegs r;
#pragma unroll
  for (int i = 0; i < SIZE; ++i) { r.c2[i][0] = g.sc2[3 * i]; r.c2[i][1] = g.sc2[3 * i + 1]; r.c2[i][2] = g.sc2[3 * i + 2]; }
  gjk_subterms<SIZE>(r, v0, v1, v2);
  const int sel = gjk_default_select<SIZE>(r);
  if (sel == 0) return false;
  // delta[s] of the subsets the scan reached goes to the owning lanes, as the general step records it
#pragma unroll
  for (int s = 1; s < (1 << SIZE); ++s) {
    const int owner = gjk_lane_of(s, gjk_elt(s, 0));
    if (lane == owner && s <= sel) g.dsum = r.dsum[s];
  }
  double ds = 1.0, num[4] = {1.0, 1.0, 1.0, 1.0};
  switch (sel) {
    case 1: gjk_pick_subset<1>(r, ds, num); break;    case 2: gjk_pick_subset<2>(r, ds, num); break;
    case 3: gjk_pick_subset<3>(r, ds, num); break;    case 4: gjk_pick_subset<4>(r, ds, num); break;
    case 5: gjk_pick_subset<5>(r, ds, num); break;    case 6: gjk_pick_subset<6>(r, ds, num); break;
    case 7: gjk_pick_subset<7>(r, ds, num); break;
    default:
      if constexpr (SIZE == 4) {
        switch (sel) {
          case 8: gjk_pick_subset<8>(r, ds, num); break;     case 9: gjk_pick_subset<9>(r, ds, num); break;
          case 10: gjk_pick_subset<10>(r, ds, num); break;   case 11: gjk_pick_subset<11>(r, ds, num); break;
          case 12: gjk_pick_subset<12>(r, ds, num); break;   case 13: gjk_pick_subset<13>(r, ds, num); break;
          case 14: gjk_pick_subset<14>(r, ds, num); break;   default: gjk_pick_subset<15>(r, ds, num); break;
        }
      }
      break;
  }
  // GJK.c:708-736: the simplex becomes subset `sel` (elements move down in order), lambdas = delta_values / delta
  __syncwarp();  // every lane has read the simplex points
  const int cs = __popc(sel);
  int rem = sel;
#pragma unroll
  for (int jj = 0; jj < SIZE; ++jj) {
    if (jj < cs) {
      const int oldpos = __ffs(rem) - 1;
      rem &= rem - 1;
      if (lane < 3 && oldpos != jj) g.sc2[3 * jj + lane] = g.sc2[3 * oldpos + lane];  // sources lie at or above jj: in-place is safe
      g.lam[jj] = num[jj] / ds;
    }
  }
  g.npts = cs;
  __syncwarp();
  return true;
}

#ifdef LQRO_PAIR_PROF  // profiling builds: cycles per part of the GJK loop (lane 0), see tools/pairprof.py
#define GJK_PROF_DECL long long gtick = clock64();
#define GJK_PROF(q) { const long long tn_ = clock64(); if (gprof) gprof[q] += tn_ - gtick; gtick = tn_; }
#define GJK_PROF_COUNT(q) { if (gprof) gprof[q] += 1; }
#else
#define GJK_PROF_DECL
#define GJK_PROF(q)
#define GJK_PROF_COUNT(q)
#endif

// Warp-cooperative GJK with the lane-parallel sub-algorithm.  scratch: LQRO_GJK_WARP_DOUBLES doubles of shared
// memory private to the warp.  Same contract as gjk_warp below.
__device__ __forceinline__ double gjk_warp_lanes(const double v0, const double v1, const double v2, const CloudView& cl, int n, double* scratch,
                                                 double* w1out, double* w2out, int* iters, long long* gprof = nullptr) {
  const int lane = lane_id();
  GjkLane g;
  g.sc2 = scratch;
  g.sprod = scratch + 12;
  {
    const int e = lane >> 3, idx = lane & 7;
    g.my_j = e;
    g.my_s = (idx & ((1 << e) - 1)) | (1 << e) | ((idx >> e) << (e + 1));
#pragma unroll
    for (int size = 2; size <= 4; ++size) {
      unsigned m = 0u;
#pragma unroll
      for (int ne = 0; ne < 4; ++ne)
        if (ne < size && !((g.my_s >> ne) & 1)) m |= 1u << gjk_lane_of(g.my_s | (1 << ne), ne);
      g.sup[size - 2] = m;
    }
  }
  g.npts = 1;  // GJK.c:340-347
  g.lam[0] = 1.0; g.lam[1] = g.lam[2] = g.lam[3] = 0.0;
  g.del = 1.0; g.dsum = 0.0;
  g.de[0] = g.de[1] = g.de[2] = g.de[3] = 0.0;
  {
    double x, y, z;
    cl.get(0, x, y, z);
    __syncwarp();
    if (lane == 0) { g.sc2[0] = x; g.sc2[1] = y; g.sc2[2] = z; }
    __syncwarp();
  }
  g.use_default = 1;
  int first_iteration = 1;
  double oldsqrd = 0.0, sqrd = 0.0;
  int max_iterations = n;  // NumVertices(obj1) * NumVertices(obj2), GJK.c:367
  int ntests = 0;
  bool exhausted = true;
  double w1[3] = {0, 0, 0}, w2[3] = {0, 0, 0};
  GJK_PROF_DECL
  while (max_iterations-- > 0) {
    GJK_PROF(5)
    const int size_in = g.npts;
    if (g.npts == 1) g.lam[0] = 1.0;
    else if (!(LQRO_GJK_STEP2 && g.npts == 2 && g.use_default && gjkl_default_step2(g, lane, v0, v1, v2)) &&
             !(LQRO_GJK_STEP3 && g.npts == 3 && g.use_default && gjkl_default_step_regs<3>(g, lane, v0, v1, v2)) &&
             !(LQRO_GJK_STEP4 && g.npts == 4 && g.use_default && gjkl_default_step_regs<4>(g, lane, v0, v1, v2)))
      gjkl_simplex_step(g, lane, v0, v1, v2);
    if (size_in == 2) { GJK_PROF(0) } else if (size_in == 3) { GJK_PROF(1) GJK_PROF_COUNT(6) } else if (size_in == 4) { GJK_PROF(2) GJK_PROF_COUNT(7) }
    (void)size_in;
    // compute_point (GJK.c:851-862), accumulated from 0; one unrolled body per simplex size (uniform branch)
    switch (g.npts) {
      case 1: gjkl_witness<1>(g, v0, v1, v2, w1, w2); break;
      case 2: gjkl_witness<2>(g, v0, v1, v2, w1, w2); break;
      case 3: gjkl_witness<3>(g, v0, v1, v2, w1, w2); break;
      default: gjkl_witness<4>(g, v0, v1, v2, w1, w2); break;
    }
    const double d0 = w2[0] - w1[0], d1 = w2[1] - w1[1], d2 = w2[2] - w1[2];
    sqrd = dot3(d0, d1, d2, d0, d1, d2);
    if (sqrd < LQRO_GJK_EPS) { exhausted = false; break; }  // GJK.c:419-422
    GJK_PROF(3)
    const double r0 = -d0, r1 = -d1, r2 = -d2;  // reverse_displacementv
    // support of the cloud in direction r (support_simple, GJK.c:770-794)
    double best = -1.7976931348623157e308;
    int besti = 0x7fffffff;
    {
      const int ns = n < cl.cap ? n : cl.cap;  // shared part, then (rarely) the overflow part, ascending index
#pragma unroll 2
      for (int m = lane; m < ns; m += 32) {
        const double v = dot3(cl.sx[m], cl.sy[m], cl.sz[m], r0, r1, r2);
        if (v > best) { best = v; besti = m; }
      }
      if (n > cl.cap) {
        int m = lane;
        while (m < cl.cap) m += 32;
        for (; m < n; m += 32) {
          const double v = dot3(cl.gx[m - cl.cap], cl.gy[m - cl.cap], cl.gz[m - cl.cap], r0, r1, r2);
          if (v > best) { best = v; besti = m; }
        }
      }
    }
    GJK_PROF(4)
    {  // warp arg-max, lowest index on ties: see gjk_warp
      const long long bits = __double_as_longlong(best + 0.0);
      const unsigned long long key = (unsigned long long)bits ^ ((unsigned long long)(bits >> 63) | 0x8000000000000000ull);
      const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      const unsigned mhi = __reduce_max_sync(LQRO_FULL, hi);
      const unsigned mlo = __reduce_max_sync(LQRO_FULL, hi == mhi ? lo : 0u);
      const bool win = hi == mhi && lo == mlo;
      besti = (int)__reduce_min_sync(LQRO_FULL, win ? (unsigned)besti : 0x7fffffffu);
      const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
      best = __longlong_as_double((long long)((mkey >> 63) ? (mkey ^ 0x8000000000000000ull) : ~mkey));
    }
    const double minus_minv = best;
    const int minp = besti;
    const double maxv = dot3(v0, v1, v2, d0, d1, d2);  // support of the single point
    ++ntests;
    double g_val = sqrd + maxv + minus_minv;  // GJK.c:450
    if (g_val < 0.0) g_val = 0;
    if (g_val < LQRO_GJK_EPS) { exhausted = false; break; }
    if ((first_iteration || sqrd < oldsqrd) && g.npts <= 3) {  // GJK.c:472-487
      double nx, ny, nz;
      cl.get(minp, nx, ny, nz);
      if (lane == 0) { g.sc2[3 * g.npts] = nx; g.sc2[3 * g.npts + 1] = ny; g.sc2[3 * g.npts + 2] = nz; }
      if (g.npts == 1) g.lam[1] = 0.0; else if (g.npts == 2) g.lam[2] = 0.0; else g.lam[3] = 0.0;
      g.npts++;
      __syncwarp();
      g.use_default = 1;
      oldsqrd = sqrd;
      first_iteration = 0;
      continue;
    }
    if (g.use_default) g.use_default = 0;  // GJK.c:490-496
    else { exhausted = false; break; }
  }
  GJK_PROF(5)
  if (iters) *iters = ntests;
#pragma unroll
  for (int d = 0; d < 3; ++d) { w1out[d] = w1[d]; w2out[d] = w2[d]; }
  if (exhausted) return 0.0;  // GJK.c:499
  return sqrd;
}

// Warp-cooperative GJK on a cloud of n points.  Returns d^2 (uniform across the warp); witnesses in w1
// (on {v_rel}) / w2 (on conv(cloud)); *iters = number of G-tests.
__device__ __forceinline__ double gjk_warp(const double v0, const double v1, const double v2, const CloudView& cl, int n, double* w1out,
                                           double* w2out, int* iters) {
  const int lane = lane_id();
  GjkRegs g;
  g.npts = 1;  // GJK.c:340-347
  g.lam[0] = 1.0;
  cl.get(0, g.c2[0][0], g.c2[0][1], g.c2[0][2]);
  g.use_default = 1;
  int first_iteration = 1;
  double oldsqrd = 0.0, sqrd = 0.0;
  int max_iterations = n;  // NumVertices(obj1) * NumVertices(obj2), GJK.c:367
  int ntests = 0;
  bool exhausted = true;
  while (max_iterations-- > 0) {
    if (g.npts == 1) g.lam[0] = 1.0;
    else if (g.npts == 2) gjk_simplex_step<2>(g, v0, v1, v2);
    else if (g.npts == 3) gjk_simplex_step<3>(g, v0, v1, v2);
    else gjk_simplex_step<4>(g, v0, v1, v2);
    // compute_point (GJK.c:851-862), accumulated from 0; one fully unrolled body per simplex size
    switch (g.npts) {
      case 1: gjk_witness<1>(g, v0, v1, v2); break;
      case 2: gjk_witness<2>(g, v0, v1, v2); break;
      case 3: gjk_witness<3>(g, v0, v1, v2); break;
      default: gjk_witness<4>(g, v0, v1, v2); break;
    }
    const double d0 = g.w2[0] - g.w1[0], d1 = g.w2[1] - g.w1[1], d2 = g.w2[2] - g.w1[2];
    sqrd = dot3(d0, d1, d2, d0, d1, d2);
    if (sqrd < LQRO_GJK_EPS) { exhausted = false; break; }  // GJK.c:419-422
    const double r0 = -d0, r1 = -d1, r2 = -d2;  // reverse_displacementv
    // support of the cloud in direction r (support_simple, GJK.c:770-794)
    double best = -1.7976931348623157e308;
    int besti = 0x7fffffff;
    {
      const int ns = n < cl.cap ? n : cl.cap;  // shared part, then (rarely) the overflow part, ascending index
#pragma unroll 2
      for (int m = lane; m < ns; m += 32) {
        const double v = dot3(cl.sx[m], cl.sy[m], cl.sz[m], r0, r1, r2);
        if (v > best) { best = v; besti = m; }
      }
      if (n > cl.cap) {
        int m = lane;
        while (m < cl.cap) m += 32;
        for (; m < n; m += 32) {
          const double v = dot3(cl.gx[m - cl.cap], cl.gy[m - cl.cap], cl.gz[m - cl.cap], r0, r1, r2);
          if (v > best) { best = v; besti = m; }
        }
      }
    }
    // warp arg-max, lowest index on ties (support_simple keeps the first maximum, GJK.c:779-790): the doubles
    // are mapped to order-preserving 64-bit keys (-0 folded into +0 first) and reduced with three REDUX
    // operations (high word, low word among the high-word winners, lowest index among the winners).
    {
      const long long bits = __double_as_longlong(best + 0.0);
      const unsigned long long key = (unsigned long long)bits ^ ((unsigned long long)(bits >> 63) | 0x8000000000000000ull);
      const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      const unsigned mhi = __reduce_max_sync(LQRO_FULL, hi);
      const unsigned mlo = __reduce_max_sync(LQRO_FULL, hi == mhi ? lo : 0u);
      const bool win = hi == mhi && lo == mlo;
      besti = (int)__reduce_min_sync(LQRO_FULL, win ? (unsigned)besti : 0x7fffffffu);
      const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
      best = __longlong_as_double((long long)((mkey >> 63) ? (mkey ^ 0x8000000000000000ull) : ~mkey));
    }
    const double minus_minv = best;
    const int minp = besti;
    const double maxv = dot3(v0, v1, v2, d0, d1, d2);  // support of the single point
    ++ntests;
    double g_val = sqrd + maxv + minus_minv;  // GJK.c:450
    if (g_val < 0.0) g_val = 0;
    if (g_val < LQRO_GJK_EPS) { exhausted = false; break; }
    if ((first_iteration || sqrd < oldsqrd) && g.npts <= 3) {  // GJK.c:472-487
      double nx, ny, nz;
      cl.get(minp, nx, ny, nz);
      switch (g.npts) {
        case 1: g.c2[1][0] = nx; g.c2[1][1] = ny; g.c2[1][2] = nz; g.lam[1] = 0.0; break;
        case 2: g.c2[2][0] = nx; g.c2[2][1] = ny; g.c2[2][2] = nz; g.lam[2] = 0.0; break;
        default: g.c2[3][0] = nx; g.c2[3][1] = ny; g.c2[3][2] = nz; g.lam[3] = 0.0; break;
      }
      g.npts++;
      g.use_default = 1;
      oldsqrd = sqrd;
      first_iteration = 0;
      continue;
    }
    if (g.use_default) g.use_default = 0;  // GJK.c:490-496
    else { exhausted = false; break; }
  }
  if (iters) *iters = ntests;
#pragma unroll
  for (int d = 0; d < 3; ++d) { w1out[d] = g.w1[d]; w2out[d] = g.w2[d]; }
  if (exhausted) return 0.0;  // GJK.c:499
  return sqrd;
}
