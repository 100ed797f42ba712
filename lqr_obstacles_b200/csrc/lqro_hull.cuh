// lqro_hull.cuh -- in-kernel convex hull for the "relative velocity inside the obstacle" case,
// replacing the reference's qconvex.exe spawn + text files (convexHull MAIN:809-838,
// convexHullFiles MAIN:762-806) and the 6-significant-digit text rendering of the cloud (MAIN:765-767).
//
// WHY IT REPLAYS QHULL'S ORDER.  The reference measures the escape distance of facet f from the FIRST
// vertex qhull lists for f (MAIN:799-802), taken from the un-rounded cloud while the normal comes from
// the 6-digit cloud: different vertices of one facet give distances differing by ~1e-5 relative, and
// ties/argmin follow qhull's facet order.  Both orders are artefacts of Quickhull's processing order
// (a facet lists its vertices by decreasing creation id, so the first one is the apex it was built
// from; facets are listed in creation order).  To agree with the reference to 1e-9 this kernel is a
// GPU Quickhull that processes points in qhull's order: same initial simplex (qh_maxsimplex), same
// first-fit partition (qh_partitionall), FIFO furthest-point processing (qh_nextfurthest), breadth-
// first visible set (qh_findhorizon), cone facets in visible/neighbour order (qh_makenew_simplicial)
// and the directed-walk partition of qh_findbest / qh_findbestnew.  Restated from qhull's published
// algorithm; checked facet-for-facet against oracle/hull_cpu.cpp (itself pinned on the reference's
// fixture and on scipy/qhull).  Merging / coplanar handling is not restated: such inputs raise the
// degenerate flag.
//
// Mapping: one warp per hull.  O(points) work -- plane distances during partition, extreme-point
// scans, the final nearest-facet scan -- is spread over the lanes; the list surgery that defines the
// order is done by lane 0.  The un-rounded cloud stays in shared memory, the rounded cloud, the facet
// pool and the outside lists live in a per-warp global workspace (L1/L2 resident).
#pragma once
#include "lqro_common.cuh"
#include "lqro_pair.cuh"

#define LQRO_HULL_MAXNEW 256  // cone facets per insertion handled by the per-lane visited bit set

// ---- strtod(printf("%g", x)): nearest double to the 6-significant-digit decimal rendering of x ------
__device__ __forceinline__ double pow10_exact(int k) {  // 10^k, exact for 0 <= k <= 22
  const double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                        1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  return t[k];
}
// value of |x| * 10^s as an unevaluated sum hi + lo (exact when 10^|s| is exact and the quotient case
// is avoided; for s < 0 the quotient is correctly rounded and lo carries the exact remainder sign)
__device__ __forceinline__ void scale10(double ax, int s, double& hi, double& lo) {
  if (s >= 0) {
    double p = (s <= 22) ? pow10_exact(s) : pow10_exact(22);
    hi = ax * p;
    lo = fma(ax, p, -hi);
    int rest = s - 22;
    while (rest > 0) {  // |x| < 1e-17: beyond exact powers, double-double product (error ~1e-32 relative)
      const int st = rest > 22 ? 22 : rest;
      const double p2 = pow10_exact(st);
      const double h2 = hi * p2;
      const double l2 = fma(hi, p2, -h2) + lo * p2;
      hi = h2 + l2;
      lo = l2 - (hi - h2);
      rest -= st;
    }
  } else {
    const double p = pow10_exact(-s > 22 ? 22 : -s);
    hi = ax / p;
    lo = -fma(hi, p, -ax) / p;  // remainder (ax - hi*p) / p, sign-exact
  }
}
__device__ __noinline__ double round6_dev(double x) {
  if (x == 0.0 || !(fabs(x) < 1.7976931348623157e308)) return x;
  const double ax = fabs(x);
  int e2;
  frexp(ax, &e2);
  int e = (int)floor((e2 - 1) * 0.30102999566398120);  // floor(log10) estimate, off by at most one
  double hi, lo;
  for (int it = 0; it < 3; ++it) {
    scale10(ax, 5 - e, hi, lo);
    if (hi < 100000.0 || (hi == 100000.0 && lo < 0.0)) { --e; continue; }
    if (hi > 1000000.0 || (hi == 1000000.0 && lo >= 0.0)) { ++e; continue; }
    break;
  }
  // round hi + lo to the nearest integer, ties to even (printf rounds the exact binary value)
  double r = rint(hi);
  const double diff = (hi - r) + lo;  // exact: |hi - r| <= 0.5 and lo is tiny
  if (diff > 0.5 || (diff == 0.5 && fmod(r, 2.0) != 0.0)) r += 1.0;
  else if (diff < -0.5 || (diff == -0.5 && fmod(r, 2.0) != 0.0)) r -= 1.0;
  const int s = 5 - e;
  double v;
  if (s >= 0) {
    if (s <= 22) v = r / pow10_exact(s);  // both exact -> correctly rounded quotient == strtod
    else { v = r / pow10_exact(22); int rest = s - 22; while (rest > 0) { const int st = rest > 22 ? 22 : rest; v /= pow10_exact(st); rest -= st; } }
  } else {
    v = r * pow10_exact(-s > 22 ? 22 : -s);  // exact product for |x| < 1e20
  }
  return x < 0 ? -v : v;
}

__global__ void k_round6(const double* __restrict__ in, int n, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = round6_dev(in[t]);
}

// ---- workspace ---------------------------------------------------------------------------------
// All indices are 16-bit (H*P <= 8192 points, <= 2*H*P+64 facets): the hot structures of one hull then fit
// in shared memory next to the tables (fast path); a hull that needs more facets than the shared pool
// holds is re-run on the full-size pool in global memory (same code, other pointers).
typedef short hidx;
struct HFacet {           // 88 bytes
  double n[3], off;
  double furthestdist;
  hidx v[3];              // vertex ids, decreasing
  hidx nb[3];             // neighbour opposite v[i]
  hidx prev, next;
  hidx ohead, olast, oprev;  // outside list (furthest last)
  hidx tag;               // scratch: BFS visit id (negative) / cone-local index
  hidx replace;
  hidx ocount;            // points in the outside list
  hidx toporient, visible, isnew, pad;   // shorts, not chars: a char store would alias every other access
};

struct HullWs {
  double *hx, *hy, *hz;   // hull input coordinates [cap_pts]            (with the hot structures)
  double *ux, *uy, *uz;   // un-rounded cloud [cap_pts]                  (global)
  double* pdist;          // [cap_pts]
  hidx *pnext, *pbest, *plist, *pstart, *vpoint;  // [cap_pts + 8]
  HFacet* F;              // [cap_f]; F[0] is the list sentinel
  hidx *order, *newf, *vis, *freel;  // [cap_f], newf [LQRO_HULL_MAXNEW]
  int* vown;              // [cap_pts + 8] ridge matching: cone facet that last touched vertex id v (-1 = none)
  int* claim;             // [cap_f] horizon search: first (facet,k) slot that reached a facet in this insertion
  int cap_pts, cap_f;
};

__host__ __device__ inline size_t hull_align16(size_t b) { return (b + 15) & ~(size_t)15; }
// bytes of the hot structures for a facet pool of cap_f (shared memory fast path, or inside the global slot)
__host__ __device__ inline size_t hull_hot_bytes(int cap_pts, int cap_f) {
  size_t b = 0;
  b += hull_align16(sizeof(HFacet) * (size_t)cap_f);
  b += hull_align16(sizeof(double) * (size_t)cap_pts);                  // pdist
  b += hull_align16(sizeof(hidx) * 5 * ((size_t)cap_pts + 8));          // pnext pbest plist pstart vpoint
  b += hull_align16(sizeof(hidx) * (3 * (size_t)cap_f + LQRO_HULL_MAXNEW));  // order vis freel newf
  b += hull_align16(3 * sizeof(double) * (size_t)cap_pts);              // hx hy hz (hull input coordinates)
  b += hull_align16(sizeof(int) * ((size_t)cap_pts + 8 + (size_t)cap_f));   // vown, claim
  return b;
}
__host__ __device__ inline int hull_full_facets(int cap_pts) { return 2 * cap_pts + 64; }
// one global slot: 6 coordinate arrays + a full-size copy of the hot structures (fallback path)
__host__ __device__ inline size_t hull_ws_bytes(int cap_pts) {
  size_t b = hull_align16(3 * sizeof(double) * (size_t)cap_pts) + hull_hot_bytes(cap_pts, hull_full_facets(cap_pts));
  return (b + 255) & ~(size_t)255;
}
__device__ __forceinline__ void hull_carve_hot(char* base, int cap_pts, int cap_f, HullWs& w) {
  w.cap_pts = cap_pts; w.cap_f = cap_f;
  w.F = reinterpret_cast<HFacet*>(base);
  base += hull_align16(sizeof(HFacet) * (size_t)cap_f);
  w.pdist = reinterpret_cast<double*>(base);
  base += hull_align16(sizeof(double) * (size_t)cap_pts);
  hidx* ip = reinterpret_cast<hidx*>(base);
  const size_t ci = (size_t)cap_pts + 8;
  w.pnext = ip; w.pbest = ip + ci; w.plist = ip + 2 * ci; w.pstart = ip + 3 * ci; w.vpoint = ip + 4 * ci;
  base += hull_align16(sizeof(hidx) * 5 * ci);
  hidx* jp = reinterpret_cast<hidx*>(base);
  w.order = jp; w.vis = jp + cap_f; w.freel = jp + 2 * (size_t)cap_f; w.newf = jp + 3 * (size_t)cap_f;
  base += hull_align16(sizeof(hidx) * (3 * (size_t)cap_f + LQRO_HULL_MAXNEW));
  double* d = reinterpret_cast<double*>(base);
  w.hx = d; w.hy = d + cap_pts; w.hz = d + 2 * (size_t)cap_pts;
  base += hull_align16(3 * sizeof(double) * (size_t)cap_pts);
  w.vown = reinterpret_cast<int*>(base);
  w.claim = w.vown + cap_pts + 8;
}
__device__ __forceinline__ void hull_carve_coords(char* gbase, int cap_pts, HullWs& w) {  // the un-rounded cloud
  double* d = reinterpret_cast<double*>(gbase);
  w.ux = d; w.uy = d + cap_pts; w.uz = d + 2 * (size_t)cap_pts;
}
// everything in the global slot (fallback / stage kernel)
__device__ __forceinline__ void hull_ws_carve(char* gbase, int cap_pts, HullWs& w) {
  hull_carve_coords(gbase, cap_pts, w);
  hull_carve_hot(gbase + hull_align16(3 * sizeof(double) * (size_t)cap_pts), cap_pts, hull_full_facets(cap_pts), w);
}

// per-warp control block in shared memory (written by lane 0, read by all after __syncwarp)
struct HullCtl {
  double DISTround, MINvisible, MINoutside, MAXcoplanar, NEARzero1, NEARzero2;
  double interior[3];
  double apex[3];
  int facet_list, facet_next, newfacet_list;
  int nfree, pool_used, nverts;
  int nvis, nnew, nlist;
  int degenerate, fail;
  int sharp, trigger_pos;
  int nfacets_out;
  int ninserted;
  int maxpoints[6];
  int simplex[4];
  long long prof[12];  // cycles per phase (lane 0): 0 init, 1 partitionall, 2 next+horizon+cone, 3 planes+match, 4 list, 5 walk, 6 redo, 7 append, 8 delete, 9 output, 10 insertions, 11 points re-partitioned
};

__device__ __forceinline__ double hdist(const HFacet& f, double x, double y, double z) {  // qh_distplane
  return f.off + x * f.n[0] + y * f.n[1] + z * f.n[2];
}
__device__ __forceinline__ double hull_distoutside(const HullCtl& c) { return 2 * c.MINoutside; }  // qh_DISToutside, max_outside = 0
__device__ __forceinline__ double hull_searchdist(const HullCtl& c) {
  return 2 * c.DISTround + (c.MINvisible > c.MAXcoplanar ? c.MINvisible : c.MAXcoplanar);
}

// list surgery (lane 0): qh_appendfacet / qh_removefacet
__device__ __forceinline__ void hl_append(HullWs& w, HullCtl& c, int f) {
  const int tail = 0;
  if (tail == c.newfacet_list) c.newfacet_list = f;
  if (tail == c.facet_next) c.facet_next = f;
  w.F[f].prev = w.F[tail].prev;
  w.F[f].next = tail;
  if (w.F[tail].prev >= 0) w.F[w.F[tail].prev].next = f; else c.facet_list = f;
  w.F[tail].prev = f;
}
__device__ __forceinline__ void hl_remove(HullWs& w, HullCtl& c, int f) {
  const int nx = w.F[f].next, pv = w.F[f].prev;
  if (f == c.newfacet_list) c.newfacet_list = nx;
  if (f == c.facet_next) c.facet_next = nx;
  if (pv >= 0) { w.F[pv].next = nx; w.F[nx].prev = pv; }
  else { c.facet_list = nx; w.F[nx].prev = -1; }
}
__device__ __forceinline__ int hl_newfacet(HullWs& w, HullCtl& c) {
  int f;
  if (c.nfree > 0) f = w.freel[--c.nfree];
  else if (c.pool_used < w.cap_f) f = c.pool_used++;
  else { c.fail = 2; return -1; }  // pool exhausted
  HFacet& F = w.F[f];
  F.prev = F.next = -1; F.ohead = F.olast = F.oprev = -1; F.furthestdist = 0; F.tag = 0; F.replace = -1; F.ocount = 0;
  F.toporient = F.visible = F.isnew = 0;
  return f;
}
// outside-list append with qhull's "furthest stays last" rule (qh_partitionpoint)
__device__ __forceinline__ void hl_outside_add(HullWs& w, int f, int p, double d) {
  HFacet& F = w.F[f];
  F.ocount++;
  if (F.olast < 0) { F.ohead = F.olast = p; F.oprev = -1; w.pnext[p] = -1; F.furthestdist = d; }
  else if (F.furthestdist < d) { w.pnext[F.olast] = p; F.oprev = F.olast; F.olast = p; w.pnext[p] = -1; F.furthestdist = d; }
  else {  // qh_setappend2ndlast
    if (F.oprev >= 0) w.pnext[F.oprev] = p; else F.ohead = p;
    w.pnext[p] = F.olast; F.oprev = p;
  }
}

// qh_setfacetplane -> qh_sethyperplane_det (3-d) + qh_normalize2
__device__ __forceinline__ void hull_setplane(HullWs& w, int fi) {
  HFacet& f = w.F[fi];
  const int a = w.vpoint[f.v[0]], b = w.vpoint[f.v[1]], c = w.vpoint[f.v[2]];
  const double p0x = w.hx[a], p0y = w.hy[a], p0z = w.hz[a];
  const double dX1 = w.hx[b] - p0x, dY1 = w.hy[b] - p0y, dZ1 = w.hz[b] - p0z;
  const double dX2 = w.hx[c] - p0x, dY2 = w.hy[c] - p0y, dZ2 = w.hz[c] - p0z;
  const double n0 = dY2 * dZ1 - dZ2 * dY1;
  const double n1 = dX1 * dZ2 - dZ1 * dX2;
  const double n2 = dX2 * dY1 - dY2 * dX1;
  double norm = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
  if (!f.toporient) norm = -norm;
  f.n[0] = n0 / norm; f.n[1] = n1 / norm; f.n[2] = n2 / norm;
  f.off = -(p0x * f.n[0] + p0y * f.n[1] + p0z * f.n[2]);
}

__device__ __forceinline__ double hdet2(double a1, double a2, double b1, double b2) { return a1 * b2 - a2 * b1; }
__device__ __forceinline__ double hdet3(double a1, double a2, double a3, double b1, double b2, double b3, double c1, double c2, double c3) {
  return a1 * hdet2(b2, b3, c2, c3) - b1 * hdet2(a2, a3, c2, c3) + c1 * hdet2(a2, a3, b2, b3);
}
// qh_detsimplex on the first k coordinates (k = current simplex size)
__device__ __forceinline__ double hull_detsimplex(const HullWs& w, const HullCtl& c, int apex, const int* simplex, int k, bool* nearzero) {
  const double ax = w.hx[apex], ay = w.hy[apex], az = w.hz[apex];
  double r[3][3];
  for (int i = 0; i < k; ++i) {
    const int p = simplex[i];
    r[i][0] = w.hx[p] - ax; r[i][1] = w.hy[p] - ay; r[i][2] = w.hz[p] - az;
  }
  double det;
  if (k == 2) { det = hdet2(r[0][0], r[0][1], r[1][0], r[1][1]); *nearzero = fabs(det) < c.NEARzero1; }
  else { det = hdet3(r[0][0], r[0][1], r[0][2], r[1][0], r[1][1], r[1][2], r[2][0], r[2][1], r[2][2]); *nearzero = fabs(det) < c.NEARzero2; }
  return det;
}

// qh_findbesthorizon restricted to what can change a non-degenerate outcome: flood over the neighbours (new or
// old) of the best facet that are within SEARCHdist of it, keeping the best.  General (rare) case, out of line.
__device__ __noinline__ int hull_horizon_flood(const HFacet* F, double searchdist, double x, double y, double z, int start,
                                               double* bestdist, int* degenerate) {
  int seen[24];
  int nseen = 0, best = start;
  int stack[24];
  int nstack = 0;
  double minsearch = *bestdist - searchdist;
  seen[nseen++] = start;
  int facet = start;
  while (true) {
    const HFacet& Fc = F[facet];
    int nextfacet = -1;
    for (int k = 0; k < 3; ++k) {
      const int nb = Fc.nb[k];
      bool was = false;
      for (int q = 0; q < nseen; ++q) if (seen[q] == nb) was = true;
      if (was) continue;
      if (nseen < 24) seen[nseen++] = nb; else { *degenerate = 1; return best; }
      const double d = hdist(F[nb], x, y, z);
      if (d > *bestdist) {
        best = nb;
        if (d > *bestdist + searchdist) nstack = 0;
        *bestdist = d;
        minsearch = d - searchdist;
      } else if (d < minsearch) continue;
      if (nextfacet >= 0) { if (nstack < 24) stack[nstack++] = nextfacet; else { *degenerate = 1; return best; } }
      nextfacet = nb;
    }
    facet = nextfacet;
    if (facet >= 0) continue;
    if (nstack == 0) break;
    facet = stack[--nstack];
  }
  return best;
}
// Common case first: if none of the three neighbours of `start` comes within SEARCHdist of the current best,
// the flood visits nothing else and the answer is `start`.
__device__ __forceinline__ int hull_findbesthorizon(const HFacet* F, double searchdist, double x, double y, double z, int start,
                                                    double* bestdist, int* degenerate) {
  const HFacet& Fs = F[start];
  const double minsearch = *bestdist - searchdist;
  const double d0 = hdist(F[Fs.nb[0]], x, y, z), d1 = hdist(F[Fs.nb[1]], x, y, z), d2 = hdist(F[Fs.nb[2]], x, y, z);
  if (d0 < minsearch && d1 < minsearch && d2 < minsearch) return start;
  return hull_horizon_flood(F, searchdist, x, y, z, start, bestdist, degenerate);
}

// Build the hull of the n points in w.hx/hy/hz.  On return (all lanes): c.fail / c.degenerate set,
// w.order[0..c.nfacets_out) = facets in qhull list order.
__device__ __forceinline__ void hull_build(HullWs& w, HullCtl& c, int n) {
  const int lane = lane_id();
  long long tprof = clock64();
  if (lane == 0) {
    for (int q = 0; q < 12; ++q) c.prof[q] = 0;
    c.degenerate = 0; c.fail = 0; c.nfree = 0; c.pool_used = 1; c.nverts = 0; c.ninserted = 0; c.nfacets_out = 0;
    c.facet_list = 0; c.facet_next = 0; c.newfacet_list = -1;
    w.F[0].prev = -1; w.F[0].next = -1; w.F[0].isnew = 0; w.F[0].visible = 0;
    if (n < 4 || n > w.cap_pts) c.fail = 1;
  }
  __syncwarp();
  if (c.fail) return;
  for (int q = lane; q < n + 8; q += 32) w.vown[q] = -1;
  for (int q = lane; q < w.cap_f; q += 32) w.claim[q] = 0x7fffffff;
  // ---- qh_maxmin: first index attaining the max / min of each coordinate
  double mxv[3], mnv[3];
  int mxi[3], mni[3];
  for (int k = 0; k < 3; ++k) { mxv[k] = -1.7976931348623157e308; mnv[k] = 1.7976931348623157e308; mxi[k] = mni[k] = 0x7fffffff; }
  for (int i = lane; i < n; i += 32) {
    const double v[3] = {w.hx[i], w.hy[i], w.hz[i]};
    for (int k = 0; k < 3; ++k) {
      if (v[k] > mxv[k]) { mxv[k] = v[k]; mxi[k] = i; }
      if (v[k] < mnv[k]) { mnv[k] = v[k]; mni[k] = i; }
    }
  }
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double ov = __shfl_xor_sync(LQRO_FULL, mxv[k], off); int oi = __shfl_xor_sync(LQRO_FULL, mxi[k], off);
      if (ov > mxv[k] || (ov == mxv[k] && oi < mxi[k])) { mxv[k] = ov; mxi[k] = oi; }
      ov = __shfl_xor_sync(LQRO_FULL, mnv[k], off); oi = __shfl_xor_sync(LQRO_FULL, mni[k], off);
      if (ov < mnv[k] || (ov == mnv[k] && oi < mni[k])) { mnv[k] = ov; mni[k] = oi; }
    }
  }
  if (lane == 0) {
    // round-off constants: qh_maxmin, qh_distround, qh_detroundoff with 'C-0' pre-merging defaults
    double MAXabs = 0, MAXsum = 0;
    for (int k = 0; k < 3; ++k) {
      const double maxcoord = mxv[k] > -mnv[k] ? mxv[k] : -mnv[k];
      if (maxcoord > MAXabs) MAXabs = maxcoord;
      MAXsum += maxcoord;
      c.maxpoints[2 * k] = mxi[k]; c.maxpoints[2 * k + 1] = mni[k];
      if (k == 1) c.NEARzero1 = 80 * MAXsum * 2.220446049250313e-16;
      if (k == 2) c.NEARzero2 = 80 * MAXsum * 2.220446049250313e-16;
    }
    double maxdistsum = sqrt(3.0) * MAXabs;
    if (MAXsum < maxdistsum) maxdistsum = MAXsum;
    c.DISTround = 2.220446049250313e-16 * (3 * maxdistsum * 1.01 + MAXabs);
    c.MINvisible = 2 * c.DISTround;
    c.MAXcoplanar = c.MINvisible;
    c.MINoutside = 2 * c.MINvisible;
    // ---- qh_maxsimplex over the extreme points
    int ns = 0;
    {
      double maxc = -1.7976931348623157e308, minc = 1.7976931348623157e308; int maxx = -1, minx = -1;
      for (int i = 0; i < 6; ++i) {
        const double px = w.hx[c.maxpoints[i]];
        if (maxc < px) { maxc = px; maxx = c.maxpoints[i]; }
        if (minc > px) { minc = px; minx = c.maxpoints[i]; }
      }
      c.simplex[ns++] = minx;
      if (maxx != minx) c.simplex[ns++] = maxx;
      if (ns < 2) c.fail = 1;
    }
    for (int k = ns; k < 4 && !c.fail; ++k) {
      int maxpoint = -1; double maxdet = -1.7976931348623157e308; bool maxnz = false;
      for (int i = 0; i < 6; ++i) {
        const int p = c.maxpoints[i];
        bool in = false;
        for (int s = 0; s < k; ++s) if (c.simplex[s] == p) in = true;
        if (in) continue;
        bool nz; const double det = fabs(hull_detsimplex(w, c, p, c.simplex, k, &nz));
        if (det > maxdet) { maxdet = det; maxpoint = p; maxnz = nz; }
      }
      if (maxpoint < 0 || maxnz) {  // qhull searches all points (rare: flat extreme set)
        for (int p = 0; p < n; ++p) {
          bool in = false;
          for (int s = 0; s < k; ++s) if (c.simplex[s] == p) in = true;
          if (in) continue;
          bool nz; const double det = fabs(hull_detsimplex(w, c, p, c.simplex, k, &nz));
          if (det > maxdet) { maxdet = det; maxpoint = p; maxnz = nz; }
        }
      }
      if (maxpoint < 0 || maxnz) { c.fail = 1; break; }
      c.simplex[k] = maxpoint;
    }
    if (!c.fail) {
      // ---- qh_initialvertices / qh_createsimplex / qh_initialhull
      for (int i = 0; i < 4; ++i) w.vpoint[i] = c.simplex[i];
      c.nverts = 4;
      const int vd[4] = {3, 2, 1, 0};
      int fid[4];
      bool top = true;
      for (int i = 0; i < 4; ++i) {
        const int f = hl_newfacet(w, c); fid[i] = f;
        int q = 0;
        for (int j = 0; j < 4; ++j) if (j != i) w.F[f].v[q++] = vd[j];
        w.F[f].toporient = top ? 1 : 0;
        hl_append(w, c, f);
        top = !top;
      }
      for (int i = 0; i < 4; ++i) { int q = 0; for (int j = 0; j < 4; ++j) if (j != i) w.F[fid[i]].nb[q++] = fid[j]; }
      c.facet_next = c.facet_list;
      for (int k = 0; k < 3; ++k) {
        double s = 0;
        for (int j = 0; j < 4; ++j) { const int p = w.vpoint[vd[j]]; s += (k == 0 ? w.hx[p] : (k == 1 ? w.hy[p] : w.hz[p])); }
        c.interior[k] = s / 4;
      }
      hull_setplane(w, fid[0]);
      if (hdist(w.F[fid[0]], c.interior[0], c.interior[1], c.interior[2]) > 0)
        for (int i = 0; i < 4; ++i) w.F[fid[i]].toporient ^= 1;
      for (int i = 0; i < 4; ++i) hull_setplane(w, fid[i]);
      for (int i = 0; i < 4; ++i)
        if (!(hdist(w.F[fid[i]], c.interior[0], c.interior[1], c.interior[2]) < -c.DISTround)) c.degenerate = 1;
      c.simplex[0] = fid[0]; c.simplex[1] = fid[1]; c.simplex[2] = fid[2]; c.simplex[3] = fid[3];  // reuse: initial facet ids
    }
  }
  __syncwarp();
  if (c.fail) return;
  if (lane == 0) { const long long tn = clock64(); c.prof[0] += tn - tprof; tprof = tn; }
  // ---- qh_partitionall: first facet (list order) a point is outside of; distances in parallel
  {
    const double distoutside = hull_distoutside(c);
    const int v0 = w.vpoint[0], v1 = w.vpoint[1], v2 = w.vpoint[2], v3 = w.vpoint[3];
    int deg = 0;
    for (int i = lane; i < n; i += 32) {
      int bestf = -1; double bd = 0;
      if (i != v0 && i != v1 && i != v2 && i != v3) {
        const double x = w.hx[i], y = w.hy[i], z = w.hz[i];
        for (int q = 0; q < 4; ++q) {
          const double d = hdist(w.F[c.simplex[q]], x, y, z);
          if (fabs(d) < 2 * distoutside) deg = 1;
          if (d >= distoutside) { bestf = c.simplex[q]; bd = d; break; }
        }
      }
      w.pbest[i] = bestf; w.pdist[i] = bd;
    }
    deg = __any_sync(LQRO_FULL, deg);
    __syncwarp();
    if (lane == 0) {
      if (deg) c.degenerate = 1;
      // build the four lists in one pass: points in index order, every record placed at the end of its reign
      {
        int bestpoint[4] = {-1, -1, -1, -1}, tail[4] = {-1, -1, -1, -1};
        double bestdist[4] = {0, 0, 0, 0};
        const int f0 = c.simplex[0], f1 = c.simplex[1], f2 = c.simplex[2];
        for (int i = 0; i < n; ++i) {
          const int f = w.pbest[i];
          if (f < 0) continue;
          const int q = (f == f0) ? 0 : (f == f1) ? 1 : (f == f2) ? 2 : 3;
          const double d = w.pdist[i];
          int emit = -1;
          if (bestpoint[q] < 0) { bestpoint[q] = i; bestdist[q] = d; }
          else if (d > bestdist[q]) { emit = bestpoint[q]; bestpoint[q] = i; bestdist[q] = d; }
          else emit = i;
          if (emit >= 0) { if (tail[q] < 0) w.F[f].ohead = (hidx)emit; else w.pnext[tail[q]] = (hidx)emit; tail[q] = emit; }
          w.F[f].ocount++;
        }
        for (int q = 0; q < 4; ++q) {
          if (bestpoint[q] < 0) continue;
          HFacet& F = w.F[c.simplex[q]];
          if (tail[q] < 0) F.ohead = (hidx)bestpoint[q]; else w.pnext[tail[q]] = (hidx)bestpoint[q];
          F.oprev = (hidx)tail[q]; F.olast = (hidx)bestpoint[q]; w.pnext[bestpoint[q]] = -1; F.furthestdist = bestdist[q];
        }
      }
      // ---- qh_furthestnext: facet with the furthest outside point moves to the front
      int bestfacet = -1; double bestdist = -1.7976931348623157e308;
      for (int f = c.facet_list; f != 0; f = w.F[f].next)
        if (w.F[f].olast >= 0 && w.F[f].furthestdist > bestdist) { bestfacet = f; bestdist = w.F[f].furthestdist; }
      if (bestfacet >= 0 && bestfacet != c.facet_list) {
        hl_remove(w, c, bestfacet);
        const int head = c.facet_list;
        w.F[bestfacet].prev = -1; w.F[bestfacet].next = head; w.F[head].prev = bestfacet;
        c.facet_list = bestfacet;
      }
      c.facet_next = c.facet_list;
    }
    __syncwarp();
  }
  if (lane == 0) { const long long tn = clock64(); c.prof[1] += tn - tprof; tprof = tn; }
  // ---- qh_buildhull
  while (true) {
    if (lane == 0) {
      // qh_nextfurthest
      int facet;
      while ((facet = c.facet_next) != 0) {
        if (w.F[facet].olast < 0) { c.facet_next = w.F[facet].next; continue; }
        break;
      }
      c.nvis = 0; c.nnew = 0; c.nlist = 0;
      if (facet != 0) {
        HFacet& F0 = w.F[facet];
        const int furthest = F0.olast;
        if (F0.oprev >= 0) { w.pnext[F0.oprev] = -1; F0.olast = F0.oprev; } else { F0.ohead = -1; F0.olast = -1; }
        F0.ocount--;
        ++c.ninserted;
        c.apex[0] = w.hx[furthest]; c.apex[1] = w.hy[furthest]; c.apex[2] = w.hz[furthest];
        F0.visible = 1; F0.replace = -1;
        F0.tag = (hidx)(-c.ninserted);
        w.vis[0] = (hidx)facet;
        c.nvis = 1;
        w.vpoint[c.nverts] = (hidx)furthest;  // the apex gets the next vertex id
      } else {
        c.nvis = -1;  // done
      }
    }
    __syncwarp();
    if (c.nvis < 0) break;
    {
      // qh_findhorizon as a wave-parallel breadth-first search.  qhull scans the visible list in order and,
      // per facet, its neighbours k = 0,1,2; a neighbour is tested once (first occurrence) and appended to
      // the list if visible.  Processing the (facet, k) slots of one wave in index order, with duplicates
      // inside a 32-slot chunk resolved to the lowest lane, yields the same list.  (qhull also moves visible
      // facets to the end of its facet list; they are deleted before anyone looks, so they stay in place.)
      const hidx vid = (hidx)(-c.ninserted);
      const double ax = c.apex[0], ay = c.apex[1], az = c.apex[2];
      int nvis = 1, wave = 0, deg = 0;
      while (wave < nvis) {
        const int wend = nvis;
        for (int b0 = wave * 3; b0 < wend * 3; b0 += 32) {
          const int idx = b0 + lane;
          const bool valid = idx < wend * 3;
          int nb = 0;
          bool test = false;
          // a facet is tested once, by the first (facet, k) slot that reaches it: claim words only ever
          // decrease (key = insertions left << 8 | slot-in-chunk), so no reset is needed between insertions
          const int key = ((0x7fff00 - (c.ninserted << 8))) | lane;
          if (valid) {
            nb = w.F[w.vis[idx / 3]].nb[idx % 3];
            if (w.F[nb].tag != vid) atomicMin(&w.claim[nb], key);
          }
          __syncwarp();
          if (valid && w.F[nb].tag != vid) test = (w.claim[nb] == key);
          __syncwarp();
          bool isvis = false;
          if (test) {
            HFacet& NB = w.F[nb];
            NB.tag = vid;
            const double d = hdist(NB, ax, ay, az);
            if (d > c.MINvisible) { isvis = true; NB.visible = 1; NB.replace = -1; }
            else if (d > -c.MAXcoplanar) deg = 1;  // coplanar horizon: qhull would merge
          }
          const unsigned m = __ballot_sync(LQRO_FULL, isvis);
          if (isvis) w.vis[nvis + __popc(m & ((1u << lane) - 1u))] = (hidx)nb;
          nvis += __popc(m);
          __syncwarp();
        }
        wave = wend;
      }
      deg = __any_sync(LQRO_FULL, deg);
      // qh_makenewfacets / qh_makenew_simplicial: one cone facet per horizon ridge, in (visible facet, k) order
      const int apex = c.nverts;
      int nnew = 0, failcode = 0;
      // pass A: count and rank the ridges
      for (int b0 = 0; b0 < nvis * 3; b0 += 32) {
        const int idx = b0 + lane;
        bool ridge = false;
        int vf = 0, nb = 0;
        if (idx < nvis * 3) { vf = w.vis[idx / 3]; nb = w.F[vf].nb[idx % 3]; ridge = !w.F[nb].visible; }
        const unsigned m = __ballot_sync(LQRO_FULL, ridge);
        if (ridge) {
          const int r = nnew + __popc(m & ((1u << lane) - 1u));
          if (r < LQRO_HULL_MAXNEW) {
            // facet id: free list first (popped from the back), then fresh pool entries
            int f;
            if (r < c.nfree) f = w.freel[c.nfree - 1 - r];
            else f = c.pool_used + (r - c.nfree);
            if (f >= w.cap_f) failcode = 2;
            else {
              w.newf[r] = (hidx)f;
              HFacet& HZ = w.F[nb];
              int horizonskip = -1;
              for (int q = 0; q < 3; ++q) if (HZ.nb[q] == vf) { horizonskip = q; break; }
              if (horizonskip < 0) failcode = 3;
              else {
                HFacet& NF = w.F[f];
                NF.ohead = NF.olast = NF.oprev = -1; NF.furthestdist = 0; NF.replace = -1; NF.ocount = 0;
                NF.visible = 0; NF.isnew = 1;
                NF.v[0] = (hidx)apex;
                int q2 = 1;
                for (int q = 0; q < 3; ++q) if (q != horizonskip) NF.v[q2++] = HZ.v[q];
                NF.toporient = HZ.toporient ? (horizonskip & 1) : ((horizonskip & 1) ^ 1);
                NF.nb[0] = (hidx)nb; NF.nb[1] = NF.nb[2] = -1;
                NF.tag = (hidx)r;  // cone-local index
                HZ.nb[horizonskip] = (hidx)f;
                // the visible facet remembers the LAST cone facet made from its ridges (f.replace)
                bool last = true;
                for (int k2 = (idx % 3) + 1; k2 < 3; ++k2) if (!w.F[w.F[vf].nb[k2]].visible) last = false;
                if (last) w.F[vf].replace = (hidx)f;
              }
            }
          } else failcode = 3;
        }
        nnew += __popc(m);
      }
      failcode = __reduce_max_sync(LQRO_FULL, failcode);
      __syncwarp();
      // pass B: append the cone facets to the facet list in rank order
      if (!failcode) {
        const int last_old = w.F[0].prev;
        for (int r = lane; r < nnew; r += 32) {
          const int f = w.newf[r];
          w.F[f].prev = (hidx)(r == 0 ? last_old : w.newf[r - 1]);
          w.F[f].next = (hidx)(r == nnew - 1 ? 0 : w.newf[r + 1]);
        }
        __syncwarp();
      }
      if (lane == 0) {
        if (deg) c.degenerate = 1;
        if (failcode) c.fail = failcode;
        else {
          c.nverts = apex + 1;
          const int used_free = nnew < c.nfree ? nnew : c.nfree;
          c.pool_used += nnew - used_free;
          c.nfree -= used_free;
          if (nnew > 0) {
            const int f0 = w.newf[0], fl = w.newf[nnew - 1];
            const int last_old = w.F[f0].prev;
            if (last_old >= 0) w.F[last_old].next = (hidx)f0; else c.facet_list = f0;
            w.F[0].prev = (hidx)fl;
            if (c.facet_next == 0) c.facet_next = f0;
            c.newfacet_list = f0;
          } else c.newfacet_list = 0;
        }
        c.nvis = nvis; c.nnew = nnew;
      }
    }
    __syncwarp();
    if (lane == 0) { const long long tn = clock64(); c.prof[2] += tn - tprof; tprof = tn; }
    if (c.nvis < 0 || c.fail) break;
    const int nnew = c.nnew;
    // qh_makenewplanes + qh_matchnewfacets (lane per cone facet)
    for (int q = lane; q < nnew; q += 32) hull_setplane(w, w.newf[q]);
    __syncwarp();
    int deg = 0;
    // qh_matchnewfacets: the cone facets {apex, a, b} meet along the ridges {apex, w}; every ridge vertex w lies
    // on exactly two of them.  The second facet to reach vown[w] links both (the neighbour opposite v[1] shares
    // {apex, v[2]} and vice versa).  sharp test of qh_sharpnewfacets folded in (orthant of the normals).
    {
      int q0s = 0;
      if (nnew > 0) { const HFacet& f0 = w.F[w.newf[0]]; q0s = (f0.n[0] > 0 ? 1 : 0) | (f0.n[1] > 0 ? 2 : 0) | (f0.n[2] > 0 ? 4 : 0); }
      int sharp = 0;
      for (int q0 = 0; q0 < nnew; q0 += 32) {
        const int q = q0 + lane;
        if (q < nnew) {
          const int f = w.newf[q];
          HFacet& fa = w.F[f];
          if (!(hdist(fa, c.interior[0], c.interior[1], c.interior[2]) < 0)) deg = 1;
          const int sg = (fa.n[0] > 0 ? 1 : 0) | (fa.n[1] > 0 ? 2 : 0) | (fa.n[2] > 0 ? 4 : 0);
          if (sg != q0s) sharp = 1;
#pragma unroll
          for (int s = 1; s <= 2; ++s) {
            const int wv = fa.v[s];
            const int prev = atomicExch(&w.vown[wv], f);
            if (prev >= 0) {
              fa.nb[3 - s] = (hidx)prev;                       // opposite the OTHER ridge vertex
              HFacet& fp = w.F[prev];
              if (fp.v[1] == wv) fp.nb[2] = (hidx)f; else fp.nb[1] = (hidx)f;
            }
          }
        }
      }
      __syncwarp();
      for (int q0 = 0; q0 < nnew; q0 += 32) {   // reset the scratch, check every ridge found its partner
        const int q = q0 + lane;
        if (q < nnew) {
          HFacet& fa = w.F[w.newf[q]];
          w.vown[fa.v[1]] = -1; w.vown[fa.v[2]] = -1;
          if (fa.nb[1] < 0 || fa.nb[2] < 0) deg = 2;
        }
      }
      sharp = __any_sync(LQRO_FULL, sharp);
      if (lane == 0) c.sharp = sharp;
    }
    deg = __reduce_max_sync(LQRO_FULL, deg);
    __syncwarp();
    if (lane == 0) { const long long tn = clock64(); c.prof[3] += tn - tprof; tprof = tn; }
    if (lane == 0) {
      if (deg == 1) c.degenerate = 1;
      if (deg == 2) c.fail = 3;
      c.trigger_pos = 0x7fffffff;
    }
    __syncwarp();
    // qh_partitionvisible, pass 1: the points to re-partition in qhull's processing order (visible facets in
    // list order, each one's outside list in order).  One lane per visible facet walks its own list; the
    // offsets come from a warp scan of the list lengths.
    {
      int nl = 0;
      const int nvis = c.nvis;
      for (int v0 = 0; v0 < nvis; v0 += 32) {
        const int vi = v0 + lane;
        int cnt = 0, vf = 0;
        if (vi < nvis) { vf = w.vis[vi]; cnt = w.F[vf].ocount; }
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(LQRO_FULL, incl, off);
          if (lane >= off) incl += t;
        }
        int pos = nl + incl - cnt;
        if (vi < nvis && cnt > 0) {
          HFacet& VF = w.F[vf];
          const hidx start = (hidx)(VF.replace >= 0 ? VF.replace : c.newfacet_list);
          for (int p = VF.ohead; p >= 0; p = w.pnext[p]) { w.plist[pos] = (hidx)p; w.pstart[pos] = start; ++pos; }
          VF.ohead = VF.olast = VF.oprev = -1; VF.ocount = 0;
        }
        nl += __shfl_sync(LQRO_FULL, incl, 31);
      }
      if (lane == 0) c.nlist = nl;
    }
    __syncwarp();
    if (lane == 0) { const long long tn = clock64(); c.prof[4] += tn - tprof; tprof = tn; c.prof[10] += 1; c.prof[11] += c.nlist; }
    if (c.fail) break;
    const int nl = c.nlist;
    if (nl > 0 && nnew > 0) {
      // pass 2 (parallel): qh_findbest(point, start, !bestoutside, isnewfacets) = directed walk over the cone
      // facets from the start facet.  Outcome per point, kept in pbest/pdist and a state in pnext (free while the
      // point is off every list): 0 = outside facet found, 1 = walk ended in [-DISTround, MINoutside): owes the
      // horizon test, 2 = walk ended below -DISTround (qhull then looks at the cone: sharp -> exhaustive search).
      int mytrig = 0x7fffffff, ldeg = 0;
      const double MINout = c.MINoutside, DISTr = c.DISTround, distoutside = hull_distoutside(c), searchdist = hull_searchdist(c);
      for (int q = lane; q < nl; q += 32) {
        const int p = w.plist[q];
        const double x = w.hx[p], y = w.hy[p], z = w.hz[p];
        int start = w.pstart[q];
        unsigned long long vm0 = 0ull, vm1 = 0ull, vm2 = 0ull, vm3 = 0ull;  // visited cone facets (<= 256)
        double bestdist = hdist(w.F[start], x, y, z);
        int best = start, state = 0;
        if (!(bestdist >= MINout)) {
          { const int t = w.F[start].tag; const unsigned long long bit = 1ull << (t & 63);
            if (t < 64) vm0 |= bit; else if (t < 128) vm1 |= bit; else if (t < 192) vm2 |= bit; else vm3 |= bit; }
          int facet = start;
          state = 1;
          while (facet >= 0) {
            const HFacet& Fc = w.F[facet];
            int nextf = -1;
            for (int k = 0; k < 3; ++k) {
              const int nb = Fc.nb[k];
              const HFacet& NB = w.F[nb];
              if (!NB.isnew) continue;
              const int t = NB.tag;
              const unsigned long long bit = 1ull << (t & 63);
              const unsigned long long cur = (t < 64) ? vm0 : (t < 128) ? vm1 : (t < 192) ? vm2 : vm3;
              if (cur & bit) continue;
              if (t < 64) vm0 |= bit; else if (t < 128) vm1 |= bit; else if (t < 192) vm2 |= bit; else vm3 |= bit;
              const double d = hdist(NB, x, y, z);
              if (d > bestdist) {
                best = nb; bestdist = d;
                if (d >= MINout) { state = 0; nextf = -1; } else nextf = nb;
                break;
              }
            }
            facet = nextf;
          }
          if (state == 1 && bestdist < -DISTr) state = 2;
        }
        w.pbest[q] = (hidx)best;
        w.pdist[q] = bestdist;
        w.pnext[p] = (hidx)state;
        if (state == 2 && q < mytrig) mytrig = q;
      }
      mytrig = (int)__reduce_min_sync(LQRO_FULL, (unsigned)mytrig);
      __syncwarp();
      if (lane == 0) { const long long tn = clock64(); c.prof[5] += tn - tprof; tprof = tn; }
      // pass 3a: sharp cone -> from the first state-2 point on (processing order) qhull searches exhaustively
      // (qh_findbestnew: cone facets in list order, cyclically from the start facet; first one at or above
      // DISToutside wins, else the best) and stays in that mode for the rest of this partition.
      const bool sharp = c.sharp != 0;
      if (mytrig != 0x7fffffff && sharp) {
        for (int q = mytrig + lane; q < nl; q += 32) {
          const int p = w.plist[q];
          const double x = w.hx[p], y = w.hy[p], z = w.hz[p];
          const int start = (q == mytrig) ? (int)w.pbest[q] : (int)w.pstart[q];
          const int s0 = w.F[start].isnew ? (int)w.F[start].tag : 0;
          double bestdist = -8.9884656743115785e307;
          int best = start, state = 1;
          int idx = s0;
          for (int it = 0; it < nnew; ++it) {
            const int f = w.newf[idx];
            const double d = hdist(w.F[f], x, y, z);
            if (d > bestdist) {
              best = f; bestdist = d;
              if (d >= distoutside) { state = 0; break; }
            }
            if (++idx == nnew) idx = 0;
          }
          w.pbest[q] = (hidx)best;
          w.pdist[q] = bestdist;
          w.pnext[p] = (hidx)state;
        }
        __syncwarp();
      }
      // pass 3b: horizon test (qh_findbesthorizon) for every point that is not outside yet
      for (int q = lane; q < nl; q += 32) {
        const int p = w.plist[q];
        int state = w.pnext[p];
        int best = w.pbest[q];
        if (state != 0) {
          double bestdist = w.pdist[q];
          best = hull_findbesthorizon(w.F, searchdist, w.hx[p], w.hy[p], w.hz[p], best, &bestdist, &ldeg);
          w.pdist[q] = bestdist;
          if (bestdist < MINout) best = -1 - best;  // not outside: inside / coplanar, dropped
        }
        w.pbest[q] = (hidx)best;
      }
      ldeg = __any_sync(LQRO_FULL, ldeg);
      __syncwarp();
      if (lane == 0) { const long long tn = clock64(); c.prof[6] += tn - tprof; tprof = tn; }
      // pass 4: append in processing order (qh_partitionpoint).  Appends to different facets commute, so lane t
      // owns cone facet t and keeps its list head / last / second-last / furthest distance in registers while
      // all lanes sweep the points in order; the rare points that land on an old facet are appended afterwards.
      {
        int pdeg = 0, anyold = 0;
        for (int q = lane; q < nl; q += 32) {  // target code: cone index, -1 dropped, -2 - facet for an old facet
          const int b = w.pbest[q];
          if (fabs(w.pdist[q]) < 4 * c.MINoutside) pdeg = 1;
          int code = -1;
          if (b >= 0) { if (w.F[b].isnew) code = w.F[b].tag; else { code = -2 - b; anyold = 1; } }
          w.pstart[q] = (hidx)code;
        }
        pdeg = __any_sync(LQRO_FULL, pdeg || ldeg);
        anyold = __any_sync(LQRO_FULL, anyold);
        __syncwarp();
        for (int t0 = 0; t0 < nnew; t0 += 32) {
          const int mine = t0 + lane;
          int oh = -1, ol = -1, op = -1, cnt = 0;
          double fd = 0.0;
          for (int c0 = 0; c0 < nl; c0 += 32) {
            // lane q holds point c0+q of the processing order; every owner collects the lanes that target its
            // facet (ascending lane = processing order) and replays exactly those appends
            const int q = c0 + lane;
            int code_l = -9, pt_l = 0;
            double d_l = 0.0;
            if (q < nl) { code_l = w.pstart[q]; pt_l = w.plist[q]; d_l = w.pdist[q]; }
            const int cn = nl - c0 < 32 ? nl - c0 : 32;
            unsigned m = 0u;
            for (int j = 0; j < cn; ++j) m |= (unsigned)(__shfl_sync(LQRO_FULL, code_l, j) == mine) << j;
            const int steps = (int)__reduce_max_sync(LQRO_FULL, (unsigned)__popc(m));
            for (int it = 0; it < steps; ++it) {
              const int src = m ? __ffs(m) - 1 : 0;
              const int pt = __shfl_sync(LQRO_FULL, pt_l, src);
              const double d = __shfl_sync(LQRO_FULL, d_l, src);
              if (m) {
                m &= m - 1;
                ++cnt;
                if (ol < 0) { oh = ol = pt; op = -1; w.pnext[pt] = -1; fd = d; }
                else if (fd < d) { w.pnext[ol] = (hidx)pt; op = ol; ol = pt; w.pnext[pt] = -1; fd = d; }
                else { if (op >= 0) w.pnext[op] = (hidx)pt; else oh = pt; w.pnext[pt] = (hidx)ol; op = pt; }  // qh_setappend2ndlast
              }
            }
          }
          if (mine < nnew && cnt > 0) {
            HFacet& NF = w.F[w.newf[mine]];
            NF.ohead = (hidx)oh; NF.olast = (hidx)ol; NF.oprev = (hidx)op; NF.furthestdist = fd; NF.ocount = (hidx)cnt;
          }
        }
        __syncwarp();
        if (lane == 0) { const long long tn = clock64(); c.prof[7] += tn - tprof; tprof = tn; }
        if (lane == 0) {
          if (pdeg) c.degenerate = 1;
          if (anyold) {
            for (int q = 0; q < nl; ++q) {
              const int code = w.pstart[q];
              if (code >= -1) continue;
              const int b = -2 - code;
              const bool wasempty = w.F[b].olast < 0;
              hl_outside_add(w, b, w.plist[q], w.pdist[q]);
              if (wasempty) { hl_remove(w, c, b); hl_append(w, c, b); }  // qhull moves it behind qh.facet_next
            }
          }
        }
      }
      __syncwarp();
    }
    if (lane == 0) { const long long tn = clock64(); c.prof[9] += tn - tprof; tprof = tn; }
    if (lane == 0) {
      // qh_deletevisible + qh_resetlists
      for (int vi = 0; vi < c.nvis; ++vi) hl_remove(w, c, w.vis[vi]);
      c.newfacet_list = -1;
    }
    {
      const int nv = c.nvis, nf0 = c.nfree;
      for (int vi = lane; vi < nv; vi += 32) {
        const int vf = w.vis[vi];
        w.F[vf].visible = 0; w.F[vf].isnew = 0; w.F[vf].tag = 0;
        w.freel[nf0 + vi] = (hidx)vf;
      }
      for (int q = lane; q < nnew; q += 32) { w.F[w.newf[q]].isnew = 0; w.F[w.newf[q]].tag = 0; }
      __syncwarp();
      if (lane == 0) c.nfree = nf0 + nv;
    }
    __syncwarp();
    if (lane == 0) { const long long tn = clock64(); c.prof[8] += tn - tprof; tprof = tn; }
  }
  if (lane == 0 && !c.fail) {
    int nf = 0;
    for (int f = c.facet_list; f != 0; f = w.F[f].next) w.order[nf++] = f;
    c.nfacets_out = nf;
  }
  __syncwarp();
}

// Nearest facet to v_rel in qhull's facet order (MAIN:825-837, with normal = facet 0's when it wins):
// distance_f = |n_f . (v_rel - p_f)|, p_f = first listed vertex of f from the UN-ROUNDED cloud (ux,uy,uz).
__device__ __forceinline__ void hull_nearest_facet(const HullWs& w, const HullCtl& c, const double* ux, const double* uy, const double* uz,
                                          double vr0, double vr1, double vr2, double* distance, double* normal, int* best_facet) {
  const int lane = lane_id();
  double bd = 1.7976931348623157e308;
  int bi = 0x7fffffff;
  for (int q = lane; q < c.nfacets_out; q += 32) {
    const HFacet& F = w.F[w.order[q]];
    const int p = w.vpoint[F.v[0]];
    const double d = fabs(F.n[0] * (vr0 - ux[p]) + F.n[1] * (vr1 - uy[p]) + F.n[2] * (vr2 - uz[p]));
    if (d < bd) { bd = d; bi = q; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(LQRO_FULL, bd, off);
    const int oi = __shfl_xor_sync(LQRO_FULL, bi, off);
    if (ov < bd || (ov == bd && oi < bi)) { bd = ov; bi = oi; }
  }
  *distance = bd;
  *best_facet = bi;
  const HFacet& F = w.F[w.order[bi]];
  normal[0] = F.n[0]; normal[1] = F.n[1]; normal[2] = F.n[2];
}

// ---- kernels -------------------------------------------------------------------------------------
// shared-memory facet pool of the fast path (facets alive at once; typical LQR-obstacle hulls need < 700)
#define LQRO_HULL_SMEM_FACETS 512


// One hull warp: its shared-memory scratch (Translate / dx / xi / step list, hot hull structures, control block)
// and its global slot (un-rounded cloud + full-size fallback pool).
struct HullWarp {
  PairWarpSmem pw;
  HullWs w;
  HullCtl* ctl;
  char* hot;
  char* gslot;
};
// bytes of shared memory one hull warp needs after the block tables
__host__ __device__ inline size_t hull_warp_smem_bytes(int H, int P) {
  return hull_align16(sizeof(double) * (3 * (size_t)H + 32 + (H + 1) / 2 + 1)) + hull_hot_bytes(H * P, LQRO_HULL_SMEM_FACETS) +
         hull_align16(sizeof(HullCtl));
}
__host__ __device__ inline size_t hull_block_smem_bytes(int H, int P) {
  return hull_align16(sizeof(double) * pair_block_table_doubles(H, P)) + hull_warp_smem_bytes(H, P);
}
__device__ __forceinline__ void hull_warp_setup(HullWarp& hw, char* sm, char* gslot, int H, int P) {
  double* wb = reinterpret_cast<double*>(sm);
  hw.pw.tr = wb; hw.pw.dx = wb + 3 * H; hw.pw.xi = hw.pw.dx + 16; hw.pw.klist = reinterpret_cast<int*>(hw.pw.xi + 16);
  hw.hot = sm + hull_align16(sizeof(double) * (3 * (size_t)H + 32 + (H + 1) / 2 + 1));
  hw.ctl = reinterpret_cast<HullCtl*>(hw.hot + hull_hot_bytes(H * P, LQRO_HULL_SMEM_FACETS));
  hw.gslot = gslot;
  hull_carve_coords(gslot, H * P, hw.w);
  hw.pw.px = hw.w.ux; hw.pw.py = hw.w.uy; hw.pw.pz = hw.w.uz;
}

// One queued inside-case pair: regenerate the reachable cloud (same code, same bits as k_pairs), render it to
// 6 digits, hull it, pick the nearest facet, write the plane.
__device__ __forceinline__ void hull_do_pair(const LqroParams& p, const PairTables& pt, HullWarp& hw, const double* __restrict__ x,
                                             const int* __restrict__ nbr, const PairOutputs& out, int pair, int* counters,
                                             long long* prof) {
  const int lane = lane_id();
  const int cap_pts = p.HP;
  PairWarpSmem& pw = hw.pw;
  HullWs& w = hw.w;
  HullCtl* ctl = hw.ctl;
  const int il = pair / p.K;
  const int i = p.local_begin + il;
  const int j = nbr[pair];
  if (lane < 16) {
    const double a = x[(size_t)i * 16 + lane], b = x[(size_t)j * 16 + lane];
    pw.xi[lane] = a;
    pw.dx[lane] = a - b;
  }
  __syncwarp();
  const int n = warp_reachable_cloud(p, pt, pw);
  hull_carve_hot(hw.hot, cap_pts, LQRO_HULL_SMEM_FACETS, w);   // fast path: hot structures + coordinates in shared memory
  for (int pass = 0; pass < 2; ++pass) {
    for (int m = lane; m < n; m += 32) {
      if (p.hull_input == 0) { w.hx[m] = round6_dev(w.ux[m]); w.hy[m] = round6_dev(w.uy[m]); w.hz[m] = round6_dev(w.uz[m]); }
      else { w.hx[m] = w.ux[m]; w.hy[m] = w.uy[m]; w.hz[m] = w.uz[m]; }
    }
    __syncwarp();
    hull_build(w, *ctl, n);
    if (ctl->fail != 2 || pass == 1) break;
    // pool too small: once more with the full-size pool in the global slot
    hull_carve_hot(hw.gslot + hull_align16(3 * sizeof(double) * (size_t)cap_pts), cap_pts, hull_full_facets(cap_pts), w);
    if (lane == 0) atomicAdd(counters + 5, 1);
  }
  if (ctl->fail) {
    if (lane == 0) { out.status[pair] = 0x100; out.nfacets[pair] = 0; atomicAdd(counters + 3, 1); }  // no plane
    __syncwarp();
    return;
  }
  double distance, normal[3]; int bestf;
  hull_nearest_facet(w, *ctl, w.ux, w.uy, w.uz, pw.dx[3], pw.dx[4], pw.dx[5], &distance, normal, &bestf);
  if (lane == 0 && prof) {
    long long tot = 0;
    for (int q = 0; q < 12; ++q) { atomicAdd((unsigned long long*)prof + q, (unsigned long long)ctl->prof[q]); if (q < 10) tot += ctl->prof[q]; }
    atomicMax(prof + 12, tot);
    atomicMax(prof + 13, (long long)n);
    atomicMax(prof + 14, (long long)ctl->nfacets_out);
  }
  if (lane == 0) {
    write_plane(out.planes + 6 * (size_t)pair, pw.xi + 3, distance, normal[0], normal[1], normal[2], true);
    out.dist[pair] = distance;
    out.nfacets[pair] = ctl->nfacets_out;
    out.status[pair] = 2 | (ctl->degenerate ? 0x100 : 0);
    if (ctl->degenerate) atomicAdd(counters + 3, 1);
  }
  __syncwarp();
}

// Stand-alone hull kernel: drains whatever the hull warps of k_pairs left in the queue (tickets continue on the
// same counter).  One warp (= one block) per queued pair at a time.
__global__ void __launch_bounds__(32, 1) k_hull(LqroParams p, LqroTables tb, const double* __restrict__ x, const int* __restrict__ nbr,
                                             PairOutputs out, HullQueue hq, int* __restrict__ work_counter, char* __restrict__ ws_base,
                                             size_t ws_stride, int* __restrict__ counters, long long* __restrict__ prof,
                                             const int* __restrict__ agent_class /* NULL: one class */, const LqroTables* __restrict__ class_tables) {
  extern __shared__ double smem[];
  const int lane = lane_id();
  int nq = *hq.count;
  if (nq > hq.capacity) nq = hq.capacity;
  if (*work_counter + (int)blockIdx.x >= nq) return;  // nothing can be left for this block
  PairTables pt;
  load_block_tables(smem, tb, p.H, p.P);
  {
    const int H = p.H, P = p.P;
    pt.NFt = smem; pt.T = smem + 16 * 3 * H; pt.sphx = pt.T + 9 * H; pt.sphy = pt.sphx + P; pt.sphz = pt.sphy + P;
    pt.CG = pt.sphz + P; pt.prn = pt.CG + 9 * H;
  }
  HullWarp hw;
  hull_warp_setup(hw, reinterpret_cast<char*>(smem) + hull_align16(sizeof(double) * pair_block_table_doubles(p.H, p.P)),
                  ws_base + ws_stride * (size_t)blockIdx.x, p.H, p.P);
  int loaded = agent_class ? -1 : 0;
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(work_counter, 1);
    item = __shfl_sync(LQRO_FULL, item, 0);
    if (item >= nq) break;
    const int pair = hq.items[item];
    if (pair < 0) continue;
    if (agent_class) {  // heterogeneous dynamics: the obstacle of pair (i, j) uses agent i's class (MAIN:1384)
      const int c = agent_class[p.local_begin + pair / p.K];
      if (c != loaded) { __syncwarp(); load_block_tables(smem, class_tables[c], p.H, p.P); loaded = c; }
    }
    hull_do_pair(p, pt, hw, x, nbr, out, pair, counters, prof);
  }
}

// Stage kernel: hull of an arbitrary cloud (AoS pts in global memory), everything in the global slot.
// mode bit 0: apply the 6-digit rendering first.  Outputs facets in qhull order (planes nf*4, verts nf*3
// as point indices); when vrel != NULL also the nearest-facet result (escape[0] = distance,
// [1..3] = normal) and best facet.
__global__ void __launch_bounds__(32, 1) k_stage_hull(const double* __restrict__ pts, int n, int mode, char* __restrict__ ws_base, int cap_pts,
                                                   double* __restrict__ planes, int* __restrict__ verts, int maxf, int* __restrict__ info /*nf, degenerate, fail, ninserted, best*/,
                                                   const double* __restrict__ vrel, double* __restrict__ escape) {
  __shared__ HullCtl ctl;
  const int lane = lane_id();
  HullWs w;
  hull_ws_carve(ws_base, cap_pts, w);
  for (int m = lane; m < n && m < cap_pts; m += 32) {
    const double a = pts[3 * m], b = pts[3 * m + 1], c = pts[3 * m + 2];
    w.ux[m] = a; w.uy[m] = b; w.uz[m] = c;
    if (mode & 1) { w.hx[m] = round6_dev(a); w.hy[m] = round6_dev(b); w.hz[m] = round6_dev(c); }
    else { w.hx[m] = a; w.hy[m] = b; w.hz[m] = c; }
  }
  __syncwarp();
  hull_build(w, ctl, n);
  if (lane == 0) { info[0] = ctl.fail ? 0 : ctl.nfacets_out; info[1] = ctl.degenerate; info[2] = ctl.fail; info[3] = ctl.ninserted; info[4] = -1; }
  __syncwarp();
  if (ctl.fail) return;
  for (int q = lane; q < ctl.nfacets_out && q < maxf; q += 32) {
    const HFacet& F = w.F[w.order[q]];
    if (planes) { planes[4 * q] = F.n[0]; planes[4 * q + 1] = F.n[1]; planes[4 * q + 2] = F.n[2]; planes[4 * q + 3] = F.off; }
    if (verts) { verts[3 * q] = w.vpoint[F.v[0]]; verts[3 * q + 1] = w.vpoint[F.v[1]]; verts[3 * q + 2] = w.vpoint[F.v[2]]; }
  }
  if (vrel) {
    double distance, normal[3]; int bestf;
    hull_nearest_facet(w, ctl, w.ux, w.uy, w.uz, vrel[0], vrel[1], vrel[2], &distance, normal, &bestf);
    if (lane == 0) { escape[0] = distance; escape[1] = normal[0]; escape[2] = normal[1]; escape[3] = normal[2]; info[4] = bestf; }
  }
}
