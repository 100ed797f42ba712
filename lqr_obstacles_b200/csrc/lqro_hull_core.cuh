// lqro_hull_core.cuh -- the convex hull of the "relative velocity inside the obstacle" case, replacing the
// reference's qconvex.exe spawn + text files (convexHull MAIN:809-838, convexHullFiles MAIN:762-806) and the
// 6-significant-digit text rendering of the cloud (MAIN:765-767).
//
// WHY IT REPLAYS QHULL'S ORDER.  The reference measures the escape distance of facet f from the FIRST vertex qhull
// lists for f (MAIN:799-802), taken from the un-rounded cloud while the normal comes from the 6-digit cloud:
// different vertices of one facet give distances differing by ~1e-5 relative, and ties / argmin follow qhull's
// facet order.  Both are artefacts of Quickhull's processing order (a facet lists its vertices by decreasing
// creation id, so the first one is the apex it was built from; facets are listed in creation order).  To agree
// with the reference to 1e-9 this is a Quickhull that processes points in qhull's order: same initial simplex
// (qh_maxsimplex), same first-fit partition (qh_partitionall), FIFO furthest-point processing (qh_nextfurthest),
// breadth-first visible set (qh_findhorizon), cone facets in visible/neighbour order (qh_makenew_simplicial) and
// the directed-walk partition of qh_findbest / qh_findbestnew.  Restated from qhull's published algorithm; checked
// facet for facet against oracle/hull_cpu.cpp (itself pinned on the reference's fixture and on scipy/qhull).
// Merging / coplanar handling is not restated: such inputs raise the degenerate flag.
//
// MAPPING: ONE WARP'S LANES PER INSERTION (plus a helper warp for two links of the chain, see HullMail), sized by
// measurement.  On the inside pairs of the 65,536-agent bench swarm (151 clouds, 14,406 insertions) an insertion sees a median of 3 visible facets, 5 new facets and 2 points to
// re-partition (90 % of the insertions: <= 5 / 7 / 33): there is no 128-wide work to spread over a CTA, the cost of
// an insertion is the LENGTH OF ITS DEPENDENT CHAIN.  So everything that defines the order is kept warp-uniform in
// registers (pool / free-list / vertex / sequence counters, the pending-facet ring), duplicates in the horizon search
// and the ridge matching are resolved with ballots and short shuffle sweeps instead of shared-memory atomics (match.any
// measured ~350 cycles per use on sm_100a and was replaced), the facet list is not a
// linked list any more (a creation sequence number per facet gives qhull's list order; a ring of pending facet ids
// gives qh_nextfurthest in one parallel probe), the outside sets are contiguous segments of a bump-allocated array whose
// appends replay qh_partitionpoint's "furthest stays last" rule as a segmented prefix maximum, and visible / new marks are
// insertion stamps that never need clearing.  Facets are SoA records (plane 32 B, topology 16 B, outside list 16 B) read with one or two
// vector loads.
//
// DUAL COMPILATION.  This file also compiles as plain C++ on top of tests/hull_host/warp_emul.h (32 host threads in
// lockstep stand in for the lanes): tests/test_hull_host.py checks the kernel's logic facet for facet against
// oracle/hull_cpu.cpp where no GPU exists.  liblqro.so never contains that build.
#pragma once
#include <float.h>
#include <math.h>

#ifndef LQRO_HULL_HOST_EMUL
#include "lqro_common.cuh"
#endif

#define LQRO_HULL_MAXNEW 256  // cone facets per insertion (visited bit set of the directed walk)
// Cones of up to LQRO_HULL_PAIRMATCH_MAX (<= 16) facets match their ridges with two lanes per facet and match.any, larger ones
// through per-vertex owner words; cones of up to LQRO_HULL_NARROW_WALK_MAX (<= 64) facets walk with a one-word visited set.
// (Macros so that the test build can force the general branches on small clouds.)
#ifndef LQRO_HULL_PAIRMATCH_MAX
#define LQRO_HULL_PAIRMATCH_MAX 16
#endif
#ifndef LQRO_HULL_NARROW_WALK_MAX
#define LQRO_HULL_NARROW_WALK_MAX 64
#endif
// Appends into a cone of up to LQRO_HULL_SMALLKEY_MAX (<= 32) facets find their buckets with ballots over the bits of the cone-local
// facet index, larger cones (and targets outside the cone) with a shuffle sweep over the facet ids.
#ifndef LQRO_HULL_SMALLKEY_MAX
#define LQRO_HULL_SMALLKEY_MAX 32
#endif
// Two warps per hull: every LQRO_HULL_SPEC_REJECT_MOD-th insertion throws the helper's speculative horizon away and searches again
// (0: never; the test build uses 3 to exercise the rejection path, which real clouds take only when a facet gets a new furthest
// point while it waits at the head of the pending list).
#ifndef LQRO_HULL_SPEC_REJECT_MOD
#define LQRO_HULL_SPEC_REJECT_MOD 0
#endif

// ---- strtod(printf("%g", x)): nearest double to the 6-significant-digit decimal rendering of x ------
#if defined(LQRO_HULL_HOST_EMUL)
static const double k_pow10_tab[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                       1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
#else
__constant__ double k_pow10_tab[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                       1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
#endif
__device__ __forceinline__ double pow10_exact(int k) { return k_pow10_tab[k]; }  // 10^k, exact for 0 <= k <= 22
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

// ---- workspace ---------------------------------------------------------------------------------
// All indices are 16-bit (H*P <= 8192 points, <= 2*H*P+64 facets): the hot structures of one hull then fit in
// shared memory next to the tables (fast path); a hull that needs more facets than the shared pool holds is
// re-run on the full-size pool in global memory (same code, other pointers).
typedef short hidx;
struct __align__(16) HPlane { double n0, n1, n2, off; };  // outward unit normal, offset
struct __align__(16) HTopo {
  hidx v[3];   // vertex ids, decreasing (v[0] = the apex the facet was built from)
  hidx nb[3];  // neighbour opposite v[i]
  hidx born;   // (insertion that created it << 1) | toporient;  < 0: not alive
  hidx vst;    // (insertion that tested it for visibility << 1) | visible
};
struct __align__(16) HOut {   // outside set of a facet: olist[ostart .. ostart + ocount), in qhull's list order, the furthest point LAST
  double fdist;
  int ostart;
  hidx ocount, pad_;
};
struct __align__(8) HMeta {
  int seq;   // position key in qhull's facet list (creation order; a facet moved to the tail gets a new key)
  int qpos;  // ring position of its pending entry, -1 = none
};

struct HullWs {
  double *hx, *hy, *hz;   // hull input coordinates [cap_pts]
  double *ux, *uy, *uz;   // un-rounded cloud [cap_pts]                  (global)
  hidx *olist, *vpoint;   // olist [ocap = 2 cap_pts]: the outside sets, one contiguous segment per facet, bump-allocated, compacted when full; vpoint [cap_pts + 8]: vertex id -> point   (hot)
  double* pdist;          // [cap_pts]       partition scratch                                       (cold: global slot)
  hidx *plist, *pstart, *pbest;  // [cap_pts + 8]                                                    (cold)
  HPlane* pl; HTopo* tp; HOut* ot; HMeta* mt;     // [cap_f]
  hidx *aux;              // [cap_f] visible facet: the last cone facet made from its ridges (qhull's f.replace); new facet: cone-local index
  hidx *vis, *freel, *ring;  // [cap_f]; ring [ring_mask + 1]
  hidx *vis2, *hst;          // [cap_f] each: the OTHER visible list and the helper warp's private stamps (two warps per hull, see HullMail)
  hidx *visB;                // the second visible-list array as carved (vis / vis2 swap on the main warp)
  hidx *order;            // [cap_f] facets in qhull's list order, stage kernel only                    (cold)
  hidx *newf;             // [LQRO_HULL_MAXNEW]
  int* vown;              // [cap_pts + 8] ridge matching of large cones: cone facet that last touched vertex id v (-1 = none)   (cold)
  int cap_pts, cap_f, ring_mask, ocap;
};

__host__ __device__ inline size_t hull_align16(size_t b) { return (b + 15) & ~(size_t)15; }
__host__ __device__ inline int hull_ring_size(int cap_f) { int r = 64; while (r < cap_f) r <<= 1; return r; }
// bytes of the HOT structures for a facet pool of cap_f (shared memory on the fast path, or inside the global slot): what every
// insertion touches -- facet records, hull input coordinates, outside-list links, vertex map, visible / free / pending lists
__host__ __device__ inline size_t hull_hot_bytes(int cap_pts, int cap_f) {
  size_t b = 0;
  b += hull_align16(sizeof(HPlane) * (size_t)cap_f);
  b += hull_align16(sizeof(HTopo) * (size_t)cap_f);
  b += hull_align16(sizeof(HOut) * (size_t)cap_f);
  b += hull_align16(sizeof(HMeta) * (size_t)cap_f);
  b += hull_align16(3 * sizeof(double) * (size_t)cap_pts);              // hx hy hz (hull input coordinates)
  b += hull_align16(sizeof(hidx) * (3 * (size_t)cap_pts + 8));          // olist vpoint
  b += hull_align16(sizeof(hidx) * (5 * (size_t)cap_f + (size_t)hull_ring_size(cap_f) + LQRO_HULL_MAXNEW));  // aux vis freel vis2 hst ring newf
  return b;
}
// bytes of the COLD structures (always in the global slot, L1 / L2 resident): the scratch of the initial partition and of
// re-partitions of more than 32 points, the owner words of cones of more than 16 facets, the facet order of the stage kernel
__host__ __device__ inline size_t hull_cold_bytes(int cap_pts, int cap_f) {
  size_t b = 0;
  b += hull_align16(sizeof(double) * (size_t)cap_pts);                  // pdist
  b += hull_align16(sizeof(hidx) * 3 * ((size_t)cap_pts + 8));          // plist pstart pbest
  b += hull_align16(sizeof(int) * ((size_t)cap_pts + 8));               // vown
  b += hull_align16(sizeof(hidx) * (size_t)cap_f);                      // order
  return b;
}
__host__ __device__ inline int hull_full_facets(int cap_pts) { return 2 * cap_pts + 64; }
// one global slot: the un-rounded cloud + the cold structures (sized for the full pool) + a full-size copy of the hot structures
// (fallback path)
__host__ __device__ inline size_t hull_ws_bytes(int cap_pts) {
  size_t b = hull_align16(3 * sizeof(double) * (size_t)cap_pts) + hull_cold_bytes(cap_pts, hull_full_facets(cap_pts)) +
             hull_hot_bytes(cap_pts, hull_full_facets(cap_pts));
  return (b + 255) & ~(size_t)255;
}
__device__ __forceinline__ void hull_carve_hot(char* base, int cap_pts, int cap_f, HullWs& w) {
  w.cap_pts = cap_pts; w.cap_f = cap_f;
  const int rs = hull_ring_size(cap_f);
  w.ring_mask = rs - 1;
  w.pl = reinterpret_cast<HPlane*>(base); base += hull_align16(sizeof(HPlane) * (size_t)cap_f);
  w.tp = reinterpret_cast<HTopo*>(base);  base += hull_align16(sizeof(HTopo) * (size_t)cap_f);
  w.ot = reinterpret_cast<HOut*>(base);   base += hull_align16(sizeof(HOut) * (size_t)cap_f);
  w.mt = reinterpret_cast<HMeta*>(base);  base += hull_align16(sizeof(HMeta) * (size_t)cap_f);
  double* d = reinterpret_cast<double*>(base);
  w.hx = d; w.hy = d + cap_pts; w.hz = d + 2 * (size_t)cap_pts;
  base += hull_align16(3 * sizeof(double) * (size_t)cap_pts);
  hidx* ip = reinterpret_cast<hidx*>(base);
  w.ocap = 2 * cap_pts;
  w.olist = ip; w.vpoint = ip + 2 * (size_t)cap_pts;
  base += hull_align16(sizeof(hidx) * (3 * (size_t)cap_pts + 8));
  hidx* jp = reinterpret_cast<hidx*>(base);
  w.aux = jp; w.vis = jp + cap_f; w.freel = jp + 2 * (size_t)cap_f;
  w.vis2 = jp + 3 * (size_t)cap_f; w.hst = jp + 4 * (size_t)cap_f; w.visB = w.vis2;
  w.ring = jp + 5 * (size_t)cap_f; w.newf = w.ring + rs;
}
__device__ __forceinline__ void hull_carve_cold(char* base, int cap_pts, int cap_f, HullWs& w) {
  w.pdist = reinterpret_cast<double*>(base); base += hull_align16(sizeof(double) * (size_t)cap_pts);
  hidx* ip = reinterpret_cast<hidx*>(base);
  const size_t ci = (size_t)cap_pts + 8;
  w.plist = ip; w.pstart = ip + ci; w.pbest = ip + 2 * ci;
  base += hull_align16(sizeof(hidx) * 3 * ci);
  w.vown = reinterpret_cast<int*>(base); base += hull_align16(sizeof(int) * ci);
  w.order = reinterpret_cast<hidx*>(base);
}
__device__ __forceinline__ void hull_carve_coords(char* gbase, int cap_pts, HullWs& w) {  // the un-rounded cloud
  double* d = reinterpret_cast<double*>(gbase);
  w.ux = d; w.uy = d + cap_pts; w.uz = d + 2 * (size_t)cap_pts;
}
// offsets inside the global slot
__host__ __device__ inline size_t hull_slot_cold_offset(int cap_pts) { return hull_align16(3 * sizeof(double) * (size_t)cap_pts); }
__host__ __device__ inline size_t hull_slot_hot_offset(int cap_pts) { return hull_slot_cold_offset(cap_pts) + hull_cold_bytes(cap_pts, hull_full_facets(cap_pts)); }
// everything in the global slot (fallback / stage kernel)
__device__ __forceinline__ void hull_ws_carve(char* gbase, int cap_pts, HullWs& w) {
  hull_carve_coords(gbase, cap_pts, w);
  hull_carve_cold(gbase + hull_slot_cold_offset(cap_pts), cap_pts, hull_full_facets(cap_pts), w);
  hull_carve_hot(gbase + hull_slot_hot_offset(cap_pts), cap_pts, hull_full_facets(cap_pts), w);
}

// per-warp control block in shared memory: constants of the hull and its results (the running counters of the build
// live in registers, identical in every lane)
struct HullCtl {
  double DISTround, MINvisible, MINoutside, MAXcoplanar, NEARzero1, NEARzero2;
  double interior[3];
  int degenerate, fail;
  int nfacets_out, ninserted, pool_used;
  int maxpoints[6];
  int simplex[4];
  long long prof[12];  // cycles per phase (lane 0, profiling builds): 0 init, 1 partitionall, 2 next + horizon, 3 cone, 4 planes + match, 5 gather, 6 walk, 7 exhaustive + horizon re-test, 8 append, 9 ring + delete, 10 insertions, 11 points re-partitioned
};

// TWO WARPS PER HULL (k_hull).  An insertion is a chain of dependent phases -- horizon search, cone, planes + ridge matching,
// gather, walk / append, ring + delete -- and one warp alone retires a dependent instruction every ~7 cycles, so the time of a
// hull is the number of instructions on that chain.  Two links can leave it:
//   * the ridge matching of the cone of insertion i depends on nothing the main warp does next: the HELPER warp does it while the
//     main warp gathers the points to re-partition, computes the planes of the cone and deletes the visible facets;
//   * the horizon search of insertion i+1 depends only on the topology and the planes, which are final once the cone of insertion i
//     is matched -- not on the outside sets that the rest of insertion i rearranges.  The facet that will be processed next is
//     the first pending one that insertion i does not delete, its apex the last point of its outside set; unless insertion i
//     appends a new furthest point to that very facet (checked: same facet, same apex, else the main warp searches itself), the
//     helper's SPECULATIVE search, run under the main warp's partition of insertion i, is the search insertion i+1 needs.  The
//     helper writes only its own visible list (w.vis2 / w.vis alternate) and its own stamps (w.hst); the main warp stamps the
//     adopted facets itself.
// One named barrier per hull, met three times per insertion (request / matching done / horizon done).  Same results as one warp:
// both warps run the same functions on the same data, only earlier.
struct HullMail {
  int hull;                     // k_hull: 1 = a hull follows, 0 = the main warp has no more work (helper leaves)
  int cmd;                      // 1: planes + matching of insertion `ins` (nnew cone facets), then the horizon below; 0: hull finished
  int ins, nnew;
  int qhead, qtail, vvis;       // main -> helper: the pending ring as insertion `ins` left it, the stamp of its visible facets
  int next_facet, next_apex, next_qhead;  // helper -> main: the speculative next insertion (next_facet < 0: none), the ring head behind it
  int req, vis_sel;             // vis_sel: 0 -> write the list to w.vis, 1 -> to w.vis2 (the one the main warp is NOT using)
  int nvis2, degen2;            // helper -> main: result of the speculative horizon
  int mdeg;                     // helper -> main: result of the ridge matching
  int bar_id;
};
#if defined(LQRO_HULL_HOST_EMUL)
static inline void hull_pair_bar(int) { emu_block_barrier(); }
#else
__device__ __forceinline__ void hull_pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
#endif

// rare-path coverage counters of the host (emulation) build: 0 cone > 16 facets (vown matching), 1 cone > 64 (wide visited
// set), 2 exhaustive mode triggered, 3 old facet moved to the tail, 4 ring compaction, 5 horizon flood, 6 multi-chunk partition,
// 7 multi-chunk horizon wave, 8 outside-set storage compacted, 9 append to a facet that already has points (segment moved),
// 10 buckets of an append found by the shuffle sweep (targets outside the cone / cone > 32 facets), 11 gather through the cold arrays,
// 12 speculative horizon of the helper warp adopted, 13 rejected (the facet got a new furthest point), 14 no speculation possible
#if defined(LQRO_HULL_HOST_EMUL)
static long long g_hull_cover[16];
#define HULL_COVER(q) { if (lane_id() == 0) ++g_hull_cover[q]; }
#else
#define HULL_COVER(q) { }
#endif
__device__ __forceinline__ double hdist(const HPlane& f, double x, double y, double z) {  // qh_distplane
  return f.off + x * f.n0 + y * f.n1 + z * f.n2;
}

// qh_setfacetplane -> qh_sethyperplane_det (3-d) + qh_normalize2, from the point indices of v[0], v[1], v[2]
__device__ __forceinline__ HPlane hull_plane(const HullWs& w, int a, int b, int c, bool toporient) {
  const double p0x = w.hx[a], p0y = w.hy[a], p0z = w.hz[a];
  const double dX1 = w.hx[b] - p0x, dY1 = w.hy[b] - p0y, dZ1 = w.hz[b] - p0z;
  const double dX2 = w.hx[c] - p0x, dY2 = w.hy[c] - p0y, dZ2 = w.hz[c] - p0z;
  const double n0 = dY2 * dZ1 - dZ2 * dY1;
  const double n1 = dX1 * dZ2 - dZ1 * dX2;
  const double n2 = dX2 * dY1 - dY2 * dX1;
  double norm = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
  if (!toporient) norm = -norm;
  HPlane f;
  f.n0 = n0 / norm; f.n1 = n1 / norm; f.n2 = n2 / norm;
  f.off = -(p0x * f.n0 + p0y * f.n1 + p0z * f.n2);
  return f;
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
__device__ __noinline__ int hull_horizon_flood(const HPlane* pl, const HTopo* tp, double searchdist, double x, double y, double z, int start,
                                               double* bestdist, int* degenerate) {
  int seen[24];
  int nseen = 0, best = start;
  int stack[24];
  int nstack = 0;
  double minsearch = *bestdist - searchdist;
  seen[nseen++] = start;
  int facet = start;
  while (true) {
    const HTopo Fc = tp[facet];
    int nextfacet = -1;
    for (int k = 0; k < 3; ++k) {
      const int nb = Fc.nb[k];
      bool was = false;
      for (int q = 0; q < nseen; ++q) if (seen[q] == nb) was = true;
      if (was) continue;
      if (nseen < 24) seen[nseen++] = nb; else { *degenerate = 1; return best; }
      const double d = hdist(pl[nb], x, y, z);
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
__device__ __forceinline__ int hull_findbesthorizon(const HPlane* pl, const HTopo* tp, double searchdist, double x, double y, double z, int start,
                                                    double* bestdist, int* degenerate) {
  const HTopo Fs = tp[start];
  const double minsearch = *bestdist - searchdist;
  const double d0 = hdist(pl[Fs.nb[0]], x, y, z), d1 = hdist(pl[Fs.nb[1]], x, y, z), d2 = hdist(pl[Fs.nb[2]], x, y, z);
  if (d0 < minsearch && d1 < minsearch && d2 < minsearch) return start;
#if defined(LQRO_HULL_HOST_EMUL)
  ++g_hull_cover[5];
#endif
  return hull_horizon_flood(pl, tp, searchdist, x, y, z, start, bestdist, degenerate);
}

// qh_findbest(point, start, !bestoutside, isnewfacets): directed walk over the cone facets from `start`.
// state: 0 = a facet the point is outside of was found, 1 = walk ended in [-DISTround, MINoutside): owes the horizon
// test, 2 = walk ended below -DISTround (qhull then looks at the cone: sharp -> exhaustive search).
// NW = 64-bit words of the visited set (cone-local indices): 1 for cones of <= 64 facets, 4 up to 256.
template <int NW>
__device__ __forceinline__ void hull_walk(const HullWs& w, int ins, double MINout, double DISTr, double x, double y, double z, int start,
                                          int& best, double& bestdist, int& state) {
  unsigned long long vm[NW];
#pragma unroll
  for (int q = 0; q < NW; ++q) vm[q] = 0ull;
  bestdist = hdist(w.pl[start], x, y, z);
  best = start; state = 0;
  if (bestdist >= MINout) return;
  {
    const int t = w.aux[start];
    if (NW == 1) vm[0] |= 1ull << t;
    else {
#pragma unroll
      for (int q = 0; q < NW; ++q) if ((t >> 6) == q) vm[q] |= 1ull << (t & 63);
    }
  }
  int facet = start;
  state = 1;
  while (facet >= 0) {
    const HTopo Fc = w.tp[facet];
    int nextf = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int nb = Fc.nb[k];
      if ((w.tp[nb].born >> 1) != ins) continue;   // cone facets only
      const int t = w.aux[nb];
      bool seen = false;
      if (NW == 1) { seen = (vm[0] >> t) & 1ull; vm[0] |= 1ull << t; }
      else {
#pragma unroll
        for (int q = 0; q < NW; ++q) if ((t >> 6) == q) { seen = (vm[q] >> (t & 63)) & 1ull; vm[q] |= 1ull << (t & 63); }
      }
      if (seen) continue;
      const double d = hdist(w.pl[nb], x, y, z);
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

// ---- outside sets as contiguous segments ------------------------------------------------------------
// qhull keeps the outside set of a facet as a list with the furthest point last; a point joins it either at the end (new
// furthest: d > furthestdist, strictly) or in front of the last one (qh_partitionpoint / qh_setappend2ndlast, and the same rule
// in qh_partitionall).  Replaying k arrivals on a list [rest, L] therefore yields [rest, e_1 .. e_k, F]: e_j is the arrival itself
// when it set no record and the record holder it displaced when it did, F the final holder.  With the list stored as a
// segment of w.olist that is a data-parallel write once every arrival knows the running maximum over the EARLIER arrivals of its
// facet -- a segmented prefix maximum over the lanes of a bucket (pointer jumping along the previous lane of the same bucket),
// ties to the earlier arrival as the strict comparison demands.

// all live segments to the front of w.olist (through w.pbest, free outside the initial partition); facets that are being
// deleted (visible in this insertion: vst == vvis) are garbage already.  Returns the new bump pointer.  Rare and out of line: scalars
// only, so that the caller keeps its workspace pointers in registers and in the shared-memory address space.
__device__ __noinline__ int hull_olist_compact(const HTopo* tp, HOut* ot, hidx* olist, hidx* pbest, int pool_used, int vvis) {
  const int lane = lane_id();
  HULL_COVER(8)
  int off = 0;
  for (int f0 = 0; f0 < pool_used; f0 += 32) {
    const int f = f0 + lane;
    int cnt = 0;
    HOut O; O.fdist = 0; O.ostart = 0; O.ocount = 0; O.pad_ = 0;
    if (f < pool_used) {
      const HTopo T = tp[f];
      if (T.born >= 0 && T.vst != vvis) { O = ot[f]; cnt = O.ocount; }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(LQRO_FULL, incl, o); if (lane >= o) incl += t; }
    const int ns = off + incl - cnt;
    for (int q = 0; q < cnt; ++q) pbest[ns + q] = olist[O.ostart + q];
    if (cnt > 0) { O.ostart = ns; ot[f] = O; }
    off += __shfl_sync(LQRO_FULL, incl, 31);
  }
  __syncwarp();
  for (int q = lane; q < off; q += 32) olist[q] = pbest[q];
  __syncwarp();
  return off;
}

// One chunk of arrivals (lane = arrival, lane order = qhull's processing order): lane `has` brings point p at distance d to facet
// f; peers = the lanes of the chunk with the same facet (0 for lanes without an arrival).  inplace: the segment of f already has
// room behind it (initial partition); otherwise the bucket gets a fresh segment at the bump pointer and whatever f holds already
// moves there (rare: an old facet reached through the horizon test, a second chunk of a long partition).  Returns the number of
// points f held before (meaningful on the lanes with an arrival).  Every lane must call.
__device__ __forceinline__ int hull_bucket_append(const HullWs& w, bool has, int f, unsigned peers, int p, double d, bool inplace, int pool_used,
                                                  int vvis, int& obump, int& failcode) {
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  const int r = __popc(peers & lt), k = __popc(peers);
  const bool leader = has && r == 0;
  HOut S; S.fdist = 0; S.ostart = 0; S.ocount = 0; S.pad_ = 0;
  int L = -1;
  if (leader) { S = w.ot[f]; if (S.ocount > 0) L = w.olist[S.ostart + S.ocount - 1]; }
  int base = S.ostart;
  if (!inplace) {
    const int need = leader ? S.ocount + k : 0;
    int incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(LQRO_FULL, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(LQRO_FULL, incl, 31);
    if (obump + total > w.ocap) {
      obump = hull_olist_compact(w.tp, w.ot, w.olist, w.pbest, pool_used, vvis);
      if (leader) S.ostart = w.ot[f].ostart;
      if (obump + total > w.ocap) { failcode = 2; return 0; }  // cannot happen: live + moved + arriving <= 2 n
    }
    base = obump + incl - need;
    obump += total;
    if (leader && S.ocount > 1) {
      HULL_COVER(9)
      for (int q = 0; q + 1 < S.ocount; ++q) w.olist[base + q] = w.olist[S.ostart + q];
    }
  }
  // the bucket learns its leader's state
  const int ld = has ? __ffs(peers) - 1 : lane;
  const int c = __shfl_sync(LQRO_FULL, (int)S.ocount, ld);
  const double f0 = __shfl_sync(LQRO_FULL, S.fdist, ld);
  const int L0 = __shfl_sync(LQRO_FULL, L, ld);
  const int b0 = __shfl_sync(LQRO_FULL, base, ld);
  // inclusive running maximum over the bucket's lanes up to this one (pointer jumping; ties stay with the earlier lane)
  const int prev = (peers & lt) ? 31 - __clz((int)(peers & lt)) : -1;
  double md = has ? d : -1.7976931348623157e308;
  int mp = p;
  {
    const int mk = __reduce_max_sync(LQRO_FULL, k);
    int pj = prev;
    for (int span = 1; span < mk; span <<= 1) {
      const int src = pj >= 0 ? pj : lane;
      const double od = __shfl_sync(LQRO_FULL, md, src);
      const int op = __shfl_sync(LQRO_FULL, mp, src);
      const int nj = __shfl_sync(LQRO_FULL, pj, src);
      if (pj >= 0) { if (!(md > od)) { md = od; mp = op; } pj = nj; }
    }
  }
  // record before this arrival: what the facet held (earliest of all), then the earlier lanes of the bucket
  double mb = c > 0 ? f0 : -1.7976931348623157e308;
  int hb = L0;
  {
    const double ed = __shfl_sync(LQRO_FULL, md, prev >= 0 ? prev : lane);
    const int ep = __shfl_sync(LQRO_FULL, mp, prev >= 0 ? prev : lane);
    if (prev >= 0 && ed > mb) { mb = ed; hb = ep; }
  }
  // final record of the bucket
  double ma = c > 0 ? f0 : -1.7976931348623157e308;
  int ha = L0;
  {
    const int last = has ? 31 - __clz((int)peers) : lane;
    const double ad = __shfl_sync(LQRO_FULL, md, last);
    const int ap = __shfl_sync(LQRO_FULL, mp, last);
    if (has && ad > ma) { ma = ad; ha = ap; }
  }
  if (has) {
    if (c > 0 || r > 0) w.olist[b0 + c - 1 + r] = (hidx)((d > mb) ? hb : p);
    if (r == 0) {
      w.olist[b0 + c + k - 1] = (hidx)ha;
      HOut N; N.fdist = ma; N.ostart = b0; N.ocount = (hidx)(c + k); N.pad_ = 0;
      w.ot[f] = N;
    }
  }
  return c;
}

// lanes with the same small key (0 <= key < 2^nbits): ballots over the key's bits
__device__ __forceinline__ unsigned hull_peers_small(bool has, int key, int nbits) {
  unsigned peers = __ballot_sync(LQRO_FULL, has);
  for (int b = 0; b < nbits; ++b) {
    const bool bit = (key >> b) & 1;
    const unsigned B = __ballot_sync(LQRO_FULL, bit);
    peers &= bit ? B : ~B;
  }
  return has ? peers : 0u;
}

// qh_findhorizon as a wave-parallel breadth-first search.  qhull scans the visible list in order and, per facet, its
// neighbours k = 0,1,2; a neighbour is tested once (first occurrence) and appended to the list if visible.  Processing the
// (facet, k) slots of one wave in index order, with duplicates inside a 32-slot chunk resolved to the lowest lane, yields the same
// list.  On entry vis[0] = the facet being processed, stamped visible; nvis = 1.
// SPEC = false: the main warp's search -- stamps in tp[].vst ((ins << 1) | visible), visible facets get aux = -1.
// SPEC = true: the helper warp's speculative search -- stamps in its private w.hst, nothing else is written but the list.
template <bool SPEC>
__device__ __forceinline__ void hull_horizon(const HullWs& w, hidx* vis, int ins, double ax, double ay, double az, double MINvisible,
                                             double MAXcoplanar, int& nvis, int& degen) {
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  const hidx vtest = (hidx)(ins << 1), vvis = (hidx)((ins << 1) | 1);
  int wave = 0;
  while (wave < nvis) {
    const int wend = nvis;
    for (int b0 = wave * 3; b0 < wend * 3; b0 += 32) {
      if (b0 > wave * 3) HULL_COVER(7)
      const int idx = b0 + lane;
      int nb = -1 - lane;
      bool fresh = false;
      if (idx < wend * 3) {
        const int vf = vis[idx / 3];
        const int cand = w.tp[vf].nb[idx % 3];
        const int stamp = SPEC ? (int)w.hst[cand] : (int)w.tp[cand].vst;
        if ((stamp >> 1) != ins) { fresh = true; nb = cand; }
      }
      // every fresh slot tests its candidate (a facet reached by several slots of the chunk is tested by each of them: same
      // result, same stamp); only a VISIBLE facet must enter the list once, through its lowest slot
      bool isvis = false;
      if (fresh) {
        const double d = hdist(w.pl[nb], ax, ay, az);
        if (d > MINvisible) isvis = true;
        else if (d > -MAXcoplanar) degen = 1;  // coplanar horizon: qhull would merge
      }
      unsigned m = __ballot_sync(LQRO_FULL, isvis);  // every lane has read the stamps before any lane writes one
      if (fresh) {
        if (SPEC) w.hst[nb] = isvis ? vvis : vtest;
        else {
          w.tp[nb].vst = isvis ? vvis : vtest;
          if (isvis) w.aux[nb] = -1;
        }
      }
      if (m & (m - 1)) {  // two or more visible candidates: drop the later slots of a repeated facet
        for (unsigned rest = m; rest; rest &= rest - 1) {
          const int j = __ffs(rest) - 1;
          const int v = __shfl_sync(LQRO_FULL, nb, j);
          if (lane > j && v == nb) isvis = false;
        }
        m = __ballot_sync(LQRO_FULL, isvis);
      }
      if (isvis) vis[nvis + __popc(m & lt)] = (hidx)nb;
      nvis += __popc(m);
      __syncwarp();
    }
    wave = wend;
  }
}

// qh_makenewplanes for the nnew cone facets w.newf[0 .. nnew): hyperplanes from the vertices, the interior point must lie below
// every one of them; the sharp test of qh_sharpnewfacets is folded in (orthant of the normals).
__device__ __forceinline__ void hull_planes(const HullWs& w, int nnew, double in0, double in1, double in2, bool& sharp, int& degen) {
  const int lane = lane_id();
  sharp = false;
  int q0s = 0;
  for (int q0 = 0; q0 < nnew; q0 += 32) {
    const int q = q0 + lane;
    int sg = 0;
    if (q < nnew) {
      const int f = w.newf[q];
      const HTopo T = w.tp[f];
      const HPlane P = hull_plane(w, w.vpoint[T.v[0]], w.vpoint[T.v[1]], w.vpoint[T.v[2]], (T.born & 1) != 0);
      w.pl[f] = P;
      if (!(hdist(P, in0, in1, in2) < 0)) degen = 1;
      sg = (P.n0 > 0 ? 1 : 0) | (P.n1 > 0 ? 2 : 0) | (P.n2 > 0 ? 4 : 0);
    }
    if (q0 == 0) q0s = __shfl_sync(LQRO_FULL, sg, 0);
    if (__any_sync(LQRO_FULL, q < nnew && sg != q0s)) sharp = true;
  }
  if (__any_sync(LQRO_FULL, degen)) degen = 1;
  __syncwarp();
}

// qh_matchnewfacets.  The cone facets {apex, a, b} meet along the ridges {apex, w}; every ridge vertex w lies on exactly two of
// them.  Returns 2 when a ridge has no partner (topology failure), else 0.  Touches only the nb[1], nb[2] fields of the cone facets.
__device__ __forceinline__ int hull_match(const HullWs& w, int nnew) {
  const int lane = lane_id();
  int mdeg = 0;
  for (int q0 = 0; q0 < nnew; q0 += 32) {
    const int q = q0 + lane;
    int f = 0, v1 = -1, v2 = -1;
    if (q < nnew) {
      f = w.newf[q];
      const HTopo T = w.tp[f];
      v1 = T.v[1]; v2 = T.v[2];
    }
    if (nnew <= LQRO_HULL_PAIRMATCH_MAX) {
      // two lanes per cone facet: lane 2r + (s - 1) holds ridge vertex v[s] of facet r; the other lane with the
      // same vertex (found by a shuffle sweep over the 2 nnew lanes) is the facet across that ridge = the neighbour opposite the OTHER ridge vertex
      const int r = lane >> 1, s = 1 + (lane & 1);
      const int v1r = __shfl_sync(LQRO_FULL, v1, r), v2r = __shfl_sync(LQRO_FULL, v2, r), fr = __shfl_sync(LQRO_FULL, f, r);
      const int wv = (r < nnew) ? (s == 1 ? v1r : v2r) : (-1 - lane);
      unsigned other = 0u;
      for (int j = 0; j < 2 * nnew; ++j) {
        const int v = __shfl_sync(LQRO_FULL, wv, j);
        if (j != lane && v == wv) other |= 1u << j;
      }
      const int pf = __shfl_sync(LQRO_FULL, fr, other ? __ffs(other) - 1 : lane);
      if (r < nnew) {
        if (__popc(other) != 1) mdeg = 2;
        else w.tp[fr].nb[3 - s] = (hidx)pf;
      }
    } else if (q < nnew) {
      HULL_COVER(0)
      // large cone: the second facet to reach vown[w] links both
      HTopo* fa = &w.tp[f];
#pragma unroll
      for (int s = 1; s <= 2; ++s) {
        const int wv = s == 1 ? v1 : v2;
        const int prev = atomicExch(&w.vown[wv], f);
        if (prev >= 0) {
          fa->nb[3 - s] = (hidx)prev;
          HTopo* fp = &w.tp[prev];
          if (fp->v[1] == wv) fp->nb[2] = (hidx)f; else fp->nb[1] = (hidx)f;
        }
      }
    }
  }
  __syncwarp();
  if (nnew > LQRO_HULL_PAIRMATCH_MAX) {
    for (int q = lane; q < nnew; q += 32) {   // reset the scratch, check every ridge found its partner
      const HTopo T = w.tp[w.newf[q]];
      w.vown[T.v[1]] = -1; w.vown[T.v[2]] = -1;
      if (T.nb[1] < 0 || T.nb[2] < 0) mdeg = 2;
    }
  }
  mdeg = __reduce_max_sync(LQRO_FULL, mdeg);
  __syncwarp();
  return mdeg;
}

// The helper warp of a hull (see HullMail): serves the main warp's requests until the hull is finished.
__device__ __forceinline__ void hull_helper(const HullWs& w, const HullCtl& c, HullMail& mail) {
  const int lane = lane_id();
  const int bar = mail.bar_id;
  for (int f = lane; f < w.cap_f; f += 32) w.hst[f] = 0;  // stamps of earlier hulls must not match this one's request numbers
  __syncwarp();
#if defined(LQRO_HULL_PROF)
  // NOTE: bar.sync compiles to BAR.SYNC.DEFER_BLOCKING -- a clock read right behind it executes before the barrier releases, so the
  // interval that FOLLOWS a barrier below contains the wait at that barrier (hp_pl is the wait at (A), not work)
  long long hp_pl = 0, hp_nn = 0, hp_planes = 0, hp_hor = 0, hp_waitA = 0, hp_waitB = 0, hp_waitC = 0, hp_n = 0, t0 = clock64(), t1;
#define HELPER_TIC(v) { t1 = clock64(); v += t1 - t0; t0 = t1; }
#else
#define HELPER_TIC(v)
#endif
  while (true) {
    hull_pair_bar(bar);                       // (A) a request is in the mailbox
    HELPER_TIC(hp_waitA)
    if (mail.cmd == 0) break;
    {
      HELPER_TIC(hp_pl)
      const int mdeg = hull_match(w, mail.nnew);
      // the facet of the NEXT insertion: the first pending one that this insertion does not delete (entries of the new cone facets
      // and of moved facets come behind every entry that exists now), its apex the last point of its outside set -- read HERE,
      // before the main warp starts to append (which moves segments of w.olist)
      int nf = -1, nap = -1, nq = 0;
      {
        const int qh = mail.qhead, qt = mail.qtail, vv = mail.vvis;
        const int pos = qh + lane;
        int f = -1; bool ok = false;
        if (pos < qt) { f = w.ring[pos & w.ring_mask]; ok = (w.mt[f].qpos == pos) && w.tp[f].vst != vv; }
        const unsigned m = __ballot_sync(LQRO_FULL, ok);
        if (m) {
          const int k = __ffs(m) - 1;
          nf = __shfl_sync(LQRO_FULL, f, k);
          nq = qh + k + 1;
          const HOut O = w.ot[nf];
          nap = w.olist[O.ostart + O.ocount - 1];
        }
      }
      if (lane == 0) { mail.mdeg = mdeg; mail.next_facet = nf; mail.next_apex = nap; mail.next_qhead = nq; }
    }
    HELPER_TIC(hp_planes)
    hull_pair_bar(bar);                       // (B) ridge matching done, next insertion named
    HELPER_TIC(hp_waitB)
    const int nf = mail.next_facet;
    if (nf >= 0) {
      const int req = mail.req, ap = mail.next_apex;
      hidx* vis = mail.vis_sel ? w.vis2 : w.vis;
      if (lane == 0) { vis[0] = (hidx)nf; w.hst[nf] = (hidx)((req << 1) | 1); }
      __syncwarp();
      int nvis = 1, degen = 0;
      hull_horizon<true>(w, vis, req, w.hx[ap], w.hy[ap], w.hz[ap], c.MINvisible, c.MAXcoplanar, nvis, degen);
      if (__any_sync(LQRO_FULL, degen)) degen = 1;
      if (lane == 0) { mail.nvis2 = nvis; mail.degen2 = degen; }
    }
    HELPER_TIC(hp_hor)
    hull_pair_bar(bar);                       // (C) speculative horizon done
    HELPER_TIC(hp_waitC)
#if defined(LQRO_HULL_PROF)
    ++hp_n; hp_nn += mail.nnew;
#endif
  }
#if defined(LQRO_HULL_PROF)
  if (lane == 0 && hp_n >= 150) printf("helper: %lld requests, %lld new facets, cycles: wait at A %lld, match + next facet %lld, horizon (+ wait at B) %lld, arrival A %lld B %lld C %lld\n", hp_n, hp_nn, hp_pl, hp_planes, hp_hor, hp_waitA, hp_waitB, hp_waitC);
#endif
}

#if defined(LQRO_HULL_PROF)
#define HULL_TIC(q) { if (lane == 0) { const long long tn_ = clock64(); hprof[q] += tn_ - tprof; tprof = tn_; } }
#define HULL_COUNT(q, v) { if (lane == 0) hprof[q] += (v); }
#else
#define HULL_TIC(q)
#define HULL_COUNT(q, v)
#endif

// Build the hull of the n points in w.hx/hy/hz.  On return (all lanes, after the trailing __syncwarp): c.fail /
// c.degenerate / c.nfacets_out / c.pool_used set; the alive facets are the pool entries [0, c.pool_used) with
// tp[f].born >= 0, their order in qhull's facet list is increasing mt[f].seq.  want_order: also fill
// w.order[0..c.nfacets_out) with the facets in that order.
// COOP: a helper warp (hull_helper) waits at the hull's named barrier (see HullMail); every path out of this function leaves it
// at its request barrier with cmd = 0.
template <bool COOP>
__device__ __forceinline__ void hull_build_t(HullWs& w, HullCtl& c, HullMail* mail, int n, bool want_order) {
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
#if defined(LQRO_HULL_PROF)
  long long hprof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprof = clock64();
#endif
  if (lane == 0) {
    c.degenerate = 0; c.fail = (n < 4 || n > w.cap_pts) ? 1 : 0; c.nfacets_out = 0; c.ninserted = 0; c.pool_used = 0;
    for (int q = 0; q < 12; ++q) c.prof[q] = 0;
  }
  __syncwarp();
  if (c.fail) { if (COOP) { if (lane == 0) mail->cmd = 0; hull_pair_bar(mail->bar_id); } return; }
  for (int q = lane; q < n + 8; q += 32) w.vown[q] = -1;
  // ---- qh_maxmin: first index attaining the max / min of each coordinate
  {
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
        // ---- qh_initialvertices / qh_createsimplex / qh_initialhull: facets 0..3, facet i lacks vertex 3 - i
        for (int i = 0; i < 4; ++i) w.vpoint[i] = (hidx)c.simplex[i];
        const int vd[4] = {3, 2, 1, 0};
        bool top[4] = {true, false, true, false};
        for (int k = 0; k < 3; ++k) {
          double s = 0;
          for (int j = 0; j < 4; ++j) { const int p = w.vpoint[vd[j]]; s += (k == 0 ? w.hx[p] : (k == 1 ? w.hy[p] : w.hz[p])); }
          c.interior[k] = s / 4;
        }
        HTopo T[4];
        for (int i = 0; i < 4; ++i) {
          int q = 0;
          for (int j = 0; j < 4; ++j) if (j != i) { T[i].v[q] = (hidx)vd[j]; T[i].nb[q] = (hidx)j; ++q; }
          T[i].vst = 0;
        }
        {
          const HPlane p0 = hull_plane(w, w.vpoint[T[0].v[0]], w.vpoint[T[0].v[1]], w.vpoint[T[0].v[2]], top[0]);
          if (hdist(p0, c.interior[0], c.interior[1], c.interior[2]) > 0)
            for (int i = 0; i < 4; ++i) top[i] = !top[i];
        }
        for (int i = 0; i < 4; ++i) {
          T[i].born = (hidx)(top[i] ? 1 : 0);
          w.tp[i] = T[i];
          const HPlane pi = hull_plane(w, w.vpoint[T[i].v[0]], w.vpoint[T[i].v[1]], w.vpoint[T[i].v[2]], top[i]);
          w.pl[i] = pi;
          if (!(hdist(pi, c.interior[0], c.interior[1], c.interior[2]) < -c.DISTround)) c.degenerate = 1;
          HMeta M; M.seq = i + 1; M.qpos = -1;
          w.mt[i] = M;
          w.aux[i] = -1;
        }
      }
    }
  }
  __syncwarp();
  if (c.fail) { if (COOP) { if (lane == 0) mail->cmd = 0; hull_pair_bar(mail->bar_id); } return; }
  HULL_TIC(0)
  const double DISTround = c.DISTround, MINvisible = c.MINvisible, MINoutside = c.MINoutside, MAXcoplanar = c.MAXcoplanar;
  const double distoutside = 2 * MINoutside;  // qh_DISToutside, max_outside = 0
  const double searchdist = 2 * DISTround + (MINvisible > MAXcoplanar ? MINvisible : MAXcoplanar);
  const double in0 = c.interior[0], in1 = c.interior[1], in2 = c.interior[2];
  int degen = 0;
  // ---- qh_partitionall: first facet (list order) a point is outside of; the four outside sets are built chunk by chunk
  int obump = 0, failc = 0;
  {
    const int v0 = w.vpoint[0], v1 = w.vpoint[1], v2 = w.vpoint[2], v3 = w.vpoint[3];
    const HPlane P0 = w.pl[0], P1 = w.pl[1], P2 = w.pl[2], P3 = w.pl[3];
    int cnt0 = 0, cnt1 = 0, cnt2 = 0, cnt3 = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      int bestf = -1; double bd = 0;
      if (i < n && i != v0 && i != v1 && i != v2 && i != v3) {
        const double x = w.hx[i], y = w.hy[i], z = w.hz[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double d = hdist(q == 0 ? P0 : q == 1 ? P1 : q == 2 ? P2 : P3, x, y, z);
          if (fabs(d) < 2 * distoutside) degen = 1;
          if (d >= distoutside) { bestf = q; bd = d; break; }
        }
      }
      if (i < n) { w.pbest[i] = (hidx)bestf; w.pdist[i] = bd; }
      cnt0 += __popc(__ballot_sync(LQRO_FULL, bestf == 0)); cnt1 += __popc(__ballot_sync(LQRO_FULL, bestf == 1));
      cnt2 += __popc(__ballot_sync(LQRO_FULL, bestf == 2)); cnt3 += __popc(__ballot_sync(LQRO_FULL, bestf == 3));
    }
    if (lane < 4) {
      HOut O; O.fdist = 0; O.ocount = 0; O.pad_ = 0;
      O.ostart = lane == 0 ? 0 : lane == 1 ? cnt0 : lane == 2 ? cnt0 + cnt1 : cnt0 + cnt1 + cnt2;
      w.ot[lane] = O;
    }
    obump = cnt0 + cnt1 + cnt2 + cnt3;
    __syncwarp();
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      int bestf = -1; double bd = 0;
      if (i < n) { bestf = w.pbest[i]; bd = w.pdist[i]; }
      const bool has = bestf >= 0;
      const unsigned peers = hull_peers_small(has, bestf, 2);
      hull_bucket_append(w, has, bestf, peers, i, bd, true, 4, -1, obump, failc);
      __syncwarp();
    }
  }
  // running state, identical in every lane
  int nfree = 0, pool_used = 4, nverts = 4, ins = 0, qhead = 0, qtail = 0, seqctr = 4;
  {
    // ---- qh_furthestnext: the facet with the furthest outside point is processed first, then list order
    int bestfacet = -1; double bestdist = -1.7976931348623157e308;
    for (int f = 0; f < 4; ++f) { const HOut O = w.ot[f]; if (O.ocount > 0 && O.fdist > bestdist) { bestfacet = f; bestdist = O.fdist; } }
    if (bestfacet >= 0) {
      if (lane == 0) { w.ring[qtail] = (hidx)bestfacet; w.mt[bestfacet].qpos = qtail; }
      ++qtail;
      for (int f = 0; f < 4; ++f) {
        if (f == bestfacet || w.ot[f].ocount <= 0) continue;
        if (lane == 0) { w.ring[qtail] = (hidx)f; w.mt[f].qpos = qtail; }
        ++qtail;
      }
    }
    __syncwarp();
  }
  HULL_TIC(1)
  // ---- qh_buildhull
  // two warps per hull (COOP): hstate = where the helper is -- 0 at its request barrier (A), 1 on its way to (B), 2 to (C);
  // spec_facet / spec_apex = the insertion whose horizon it was asked to search, req the request number, vcur which visible list
  // the main warp is using (0: w.vis, 1: w.vis2 -- swapped on adoption so that w.vis is always the current one)
  int hstate = 0, spec_facet = -1, spec_apex = -1, spec_qhead = 0, req = 0;
  while (true) {
    // qh_nextfurthest: the first pending facet of the list = the first ring entry that is still its facet's entry
    int facet = -1;
    if (COOP && spec_facet >= 0) {
      // the helper named this facet when the previous insertion had fixed its visible set: every ring entry in front of it is stale
      // or belongs to a facet that insertion deleted
      facet = spec_facet; qhead = spec_qhead;
    } else {
      while (qhead < qtail) {
        const int pos = qhead + lane;
        int f = -1; bool ok = false;
        if (pos < qtail) { f = w.ring[pos & w.ring_mask]; ok = (w.mt[f].qpos == pos); }
        const unsigned m = __ballot_sync(LQRO_FULL, ok);
        if (m) { const int k = __ffs(m) - 1; facet = __shfl_sync(LQRO_FULL, f, k); qhead += k + 1; break; }
        qhead = (qhead + 32 < qtail) ? qhead + 32 : qtail;
      }
    }
    if (COOP && hstate == 2) { hull_pair_bar(mail->bar_id); hstate = 0; }   // (C) the helper's speculative horizon is complete
    if (facet < 0) break;
    ++ins;
    const hidx vvis = (hidx)((ins << 1) | 1);
    int nvis = 1;
    double ax, ay, az;
    int furthest;
    {
      const HOut O = w.ot[facet];
      furthest = w.olist[O.ostart + O.ocount - 1];
      ax = w.hx[furthest]; ay = w.hy[furthest]; az = w.hz[furthest];
    }
    const bool adopt = COOP && spec_facet == facet && spec_apex == furthest && !(LQRO_HULL_SPEC_REJECT_MOD > 0 && ins % (LQRO_HULL_SPEC_REJECT_MOD > 0 ? LQRO_HULL_SPEC_REJECT_MOD : 1) == 0);
    if (COOP) { if (adopt) HULL_COVER(12) else if (spec_facet >= 0) HULL_COVER(13) else HULL_COVER(14) }
    if (adopt) {
      // the helper searched this very horizon during the previous partition: take its list, stamp the visible facets
      nvis = mail->nvis2;
      if (mail->degen2) degen = 1;
      hidx* t = w.vis; w.vis = w.vis2; w.vis2 = t;
      for (int vi = lane; vi < nvis; vi += 32) { const int vf = w.vis[vi]; w.tp[vf].vst = vvis; w.aux[vf] = -1; }
      if (lane == 0) { w.vpoint[nverts] = (hidx)furthest; w.mt[facet].qpos = -1; }
      __syncwarp();
    } else {
      if (lane == 0) {
        w.vpoint[nverts] = (hidx)furthest;  // the apex gets the next vertex id
        w.vis[0] = (hidx)facet;
        w.tp[facet].vst = vvis;
        w.aux[facet] = -1;
        w.mt[facet].qpos = -1;
      }
      __syncwarp();
      hull_horizon<false>(w, w.vis, ins, ax, ay, az, MINvisible, MAXcoplanar, nvis, degen);
    }
    spec_facet = -1;
    HULL_TIC(2)
    // qh_makenewfacets / qh_makenew_simplicial: one cone facet per horizon ridge, in (visible facet, k) order
    const int apexv = nverts;
    int nnew = 0;
    {
      int failcode = 0;
      for (int b0 = 0; b0 < nvis * 3; b0 += 32) {
        const int idx = b0 + lane;
        bool ridge = false;
        int vf = 0, nb = 0, k = 0;
        HTopo VT;
        if (idx < nvis * 3) { vf = w.vis[idx / 3]; k = idx % 3; VT = w.tp[vf]; nb = k == 0 ? VT.nb[0] : k == 1 ? VT.nb[1] : VT.nb[2]; ridge = w.tp[nb].vst != vvis; }
        const unsigned m = __ballot_sync(LQRO_FULL, ridge);
        if (ridge) {
          const int r = nnew + __popc(m & lt);
          if (r < LQRO_HULL_MAXNEW) {
            // facet id: free list first (popped from the back), then fresh pool entries
            const int f = (r < nfree) ? (int)w.freel[nfree - 1 - r] : pool_used + (r - nfree);
            if (f >= w.cap_f) failcode = 2;
            else {
              w.newf[r] = (hidx)f;
              const HTopo HZ = w.tp[nb];
              const int hs = (HZ.nb[0] == vf) ? 0 : (HZ.nb[1] == vf) ? 1 : (HZ.nb[2] == vf) ? 2 : -1;
              if (hs < 0) failcode = 3;
              else {
                HTopo NF;
                NF.v[0] = (hidx)apexv;
                NF.v[1] = hs == 0 ? HZ.v[1] : HZ.v[0];
                NF.v[2] = hs == 2 ? HZ.v[1] : HZ.v[2];
                const int top = (HZ.born & 1) ? (hs & 1) : ((hs & 1) ^ 1);
                NF.nb[0] = (hidx)nb; NF.nb[1] = NF.nb[2] = -1;
                NF.born = (hidx)((ins << 1) | top);
                NF.vst = 0;
                w.tp[f] = NF;
                w.tp[nb].nb[hs] = (hidx)f;
                w.aux[f] = (hidx)r;  // cone-local index
                HOut O; O.fdist = 0; O.ostart = 0; O.ocount = 0; O.pad_ = 0;
                w.ot[f] = O;
                HMeta M; M.seq = seqctr + 1 + r; M.qpos = -1;
                w.mt[f] = M;
                // the visible facet remembers the LAST cone facet made from its ridges (qhull's f.replace)
                bool last = true;
                if (k < 1 && w.tp[VT.nb[1]].vst != vvis) last = false;
                if (k < 2 && w.tp[VT.nb[2]].vst != vvis) last = false;
                if (last) w.aux[vf] = (hidx)f;
              }
            }
          } else failcode = 3;
        }
        nnew += __popc(m);
      }
      failcode = __reduce_max_sync(LQRO_FULL, failcode);
      __syncwarp();
      if (failcode) { failc = failcode; break; }
      const int used_free = nnew < nfree ? nnew : nfree;
      pool_used += nnew - used_free;
      nfree -= used_free;
      nverts = apexv + 1;
      seqctr += nnew;
    }
    HULL_TIC(3)
    // qh_makenewplanes + qh_matchnewfacets (hull_planes_match): inline, or on the helper warp while this warp gathers the points
    bool sharp = false;
    if (COOP) {
      ++req;
      if (lane == 0) {
        mail->cmd = 1; mail->ins = ins; mail->nnew = nnew; mail->req = req;
        mail->qhead = qhead; mail->qtail = qtail; mail->vvis = vvis;
        mail->vis_sel = (w.vis2 == w.visB) ? 1 : 0;
      }
      hull_pair_bar(mail->bar_id);   // (A)
      hstate = 1;
    } else {
      hull_planes(w, nnew, in0, in1, in2, sharp, degen);
      HULL_TIC(4)
      if (hull_match(w, nnew) == 2) { failc = 3; break; }
      HULL_TIC(9)   // profiling builds, one warp per hull: ridge matching is charged to "ring + delete"
    }
    HULL_TIC(4)
    // qh_partitionvisible, pass 1: the points to re-partition in qhull's processing order (visible facets in list
    // order, each one's outside set in order).  Common case (<= 32 visible facets, <= 32 points): the lanes locate their point
    // in the segments of the visible facets through a warp scan of the set sizes, nothing is staged.  Otherwise the
    // points go through the cold arrays plist / pstart.
    int nl = 0;
    bool staged = true;
    int gp = 0, gs = 0;
    const hidx newfirst = w.newf[0];
    if (nvis <= 32) {
      int cnt = 0, ost = 0, sf = 0;
      if (lane < nvis) {
        const int vf = w.vis[lane];
        const HOut O = w.ot[vf];
        cnt = O.ocount - (lane == 0 ? 1 : 0);  // vis[0] is the facet being processed: its last point is the apex
        ost = O.ostart;
        const hidx rp = w.aux[vf];
        sf = rp >= 0 ? rp : newfirst;
      }
      int incl = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(LQRO_FULL, incl, off);
        if (lane >= off) incl += t;
      }
      nl = __shfl_sync(LQRO_FULL, incl, 31);
      if (nl <= 32) {
        staged = false;
        // source facet of point `lane`: the number of visible facets whose sets end at or before it
        int idx = 0;
        if (nvis <= 6) {
          for (int v = 0; v + 1 < nvis; ++v) idx += (__shfl_sync(LQRO_FULL, incl, v) <= lane) ? 1 : 0;
        } else {
#pragma unroll
          for (int st = 16; st >= 1; st >>= 1) {
            const int e = __shfl_sync(LQRO_FULL, incl, idx + st - 1);
            if (e <= lane) idx += st;
          }
          if (idx > 31) idx = 31;
        }
        const int so = __shfl_sync(LQRO_FULL, ost, idx), se = __shfl_sync(LQRO_FULL, incl - cnt, idx);
        gs = __shfl_sync(LQRO_FULL, sf, idx);
        if (lane < nl) gp = w.olist[so + (lane - se)];
      }
    }
    if (staged) {
      HULL_COVER(11)
      nl = 0;
      for (int v0 = 0; v0 < nvis; v0 += 32) {
        const int vi = v0 + lane;
        int cnt = 0, vf = 0, ost = 0;
        if (vi < nvis) { vf = w.vis[vi]; const HOut O = w.ot[vf]; cnt = O.ocount - (vi == 0 ? 1 : 0); ost = O.ostart; }
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(LQRO_FULL, incl, off);
          if (lane >= off) incl += t;
        }
        int pos = nl + incl - cnt;
        if (cnt > 0) {
          const hidx rp = w.aux[vf];
          const hidx start = rp >= 0 ? rp : newfirst;
          for (int t = 0; t < cnt; ++t) { w.plist[pos] = w.olist[ost + t]; w.pstart[pos] = start; ++pos; }
        }
        nl += __shfl_sync(LQRO_FULL, incl, 31);
      }
      __syncwarp();
    }
    if (COOP) {
      // next to the helper's ridge matching: the planes of the cone (measured: the same FP64 chain takes 3.5x longer on the helper
      // warp than here, so it stays on this warp) and qh_deletevisible (the lists of the visible facets have just been gathered)
      HULL_TIC(5)
      hull_planes(w, nnew, in0, in1, in2, sharp, degen);
      for (int vi = lane; vi < nvis; vi += 32) {
        const int vf = w.vis[vi];
        w.tp[vf].born = -1;
        w.mt[vf].qpos = -1;
        w.freel[nfree + vi] = (hidx)vf;
      }
      nfree += nvis;
      HULL_TIC(9)
      hull_pair_bar(mail->bar_id);   // (B) the cone is matched, the next insertion named
      HULL_TIC(4)                    // profiling builds: the wait for the helper is charged to "planes + match"
      hstate = 2;
      spec_facet = mail->next_facet; spec_apex = mail->next_apex; spec_qhead = mail->next_qhead;
      if (mail->mdeg == 2) { failc = 3; break; }
    }
    HULL_TIC(5)
    HULL_COUNT(10, 1)
    HULL_COUNT(11, nl)
    int nmoved = 0;
    if (nl > 0 && nnew > 0) {
      // pass 2, chunks of 32 points in processing order: directed walk; sharp cone -> from the first point whose
      // walk ended below -DISTround on, qhull searches exhaustively (qh_findbestnew: cone facets in list order,
      // cyclically from the start facet; first one at or above DISToutside wins, else the best) and stays in that
      // mode for the rest of this partition; horizon test (qh_findbesthorizon) for every point that is not outside
      // yet; append (qh_partitionpoint).
      bool exh = false;
      int ldeg = 0;
      const int cbits = nnew <= LQRO_HULL_SMALLKEY_MAX ? 32 - __clz(nnew - 1 > 0 ? nnew - 1 : 1) : 0;  // bits of a cone-local index
      for (int c0 = 0; c0 < nl; c0 += 32) {
        if (c0 > 0) HULL_COVER(6)
        const int q = c0 + lane;
        const bool valid = q < nl;
        int p = 0, start = 0, best = -1, state = 0;
        double x = 0, y = 0, z = 0, bestdist = 0;
        if (valid) {
          if (staged) { p = w.plist[q]; start = w.pstart[q]; } else { p = gp; start = gs; }
          x = w.hx[p]; y = w.hy[p]; z = w.hz[p]; best = start;
        }
        int from = exh ? 0 : 32;
        bool trigger_chunk = false;
        if (!exh) {
          if (valid) {
            if (nnew <= LQRO_HULL_NARROW_WALK_MAX) hull_walk<1>(w, ins, MINoutside, DISTround, x, y, z, start, best, bestdist, state);
            else { HULL_COVER(1) hull_walk<4>(w, ins, MINoutside, DISTround, x, y, z, start, best, bestdist, state); }
          }
          HULL_TIC(6)
          if (sharp) {
            const unsigned trig = __ballot_sync(LQRO_FULL, valid && state == 2);
            if (trig) { from = __ffs(trig) - 1; exh = true; trigger_chunk = true; HULL_COVER(2) }
          }
        }
        if (valid && lane >= from) {
          // the point that triggered the exhaustive mode starts from where its walk ended, the later ones from their start facet
          const int sf = (trigger_chunk && lane == from) ? best : start;
          const int s0 = ((w.tp[sf].born >> 1) == ins) ? (int)w.aux[sf] : 0;
          double bd = -8.9884656743115785e307;
          int b = sf, st = 1, idx = s0;
          for (int it = 0; it < nnew; ++it) {
            const int f = w.newf[idx];
            const double d = hdist(w.pl[f], x, y, z);
            if (d > bd) {
              b = f; bd = d;
              if (d >= distoutside) { st = 0; break; }
            }
            if (++idx == nnew) idx = 0;
          }
          best = b; bestdist = bd; state = st;
        }
        if (valid && state != 0) {
          best = hull_findbesthorizon(w.pl, w.tp, searchdist, x, y, z, best, &bestdist, &ldeg);
          if (bestdist < MINoutside) best = -1;  // not outside: inside / coplanar, dropped
        }
        if (valid && fabs(bestdist) < 4 * MINoutside) degen = 1;
        HULL_TIC(7)
        // append (qh_partitionpoint): appends to different facets commute; the lanes of one target facet form a bucket
        // (hull_bucket_append).  Targets inside a cone of <= 32 facets are told apart by their cone-local index.
        const bool has = valid && best >= 0;
        bool cone = false;
        int key = 0;
        if (has) { cone = (w.tp[best].born >> 1) == ins; key = w.aux[best]; }
        unsigned peers;
        if (nnew <= LQRO_HULL_SMALLKEY_MAX && !__any_sync(LQRO_FULL, has && !cone)) {
          peers = hull_peers_small(has, key, cbits);
        } else {
          HULL_COVER(10)
          peers = 0u;
          const int fk = has ? best : -1 - lane;
          const int cn = nl - c0 < 32 ? nl - c0 : 32;
          for (int j = 0; j < cn; ++j) peers |= (unsigned)(__shfl_sync(LQRO_FULL, fk, j) == fk) << j;
          if (!has) peers = 0u;
        }
        int fcode = 0;
        const int cold = hull_bucket_append(w, has, best, peers, p, bestdist, false, pool_used, vvis, obump, fcode);
        if (fcode) failc = fcode;
        // an old facet gets its first point: qhull moves it to the tail of its facet list
        const bool moved = has && (peers & lt) == 0u && cold == 0 && !cone;
        const unsigned mv = __ballot_sync(LQRO_FULL, moved);
        if (moved) w.vis[nvis + nmoved + __popc(mv & lt)] = (hidx)best;  // behind the visible list (disjoint sets of alive facets)
        nmoved += __popc(mv);
        if (mv) HULL_COVER(3)
        __syncwarp();
        HULL_TIC(8)
      }
      if (__any_sync(LQRO_FULL, ldeg)) degen = 1;
      if (failc) break;
    }
    // the pending ring: cone facets with points, in creation order, then the old facets that got their first point
    // (qhull moved those to the tail of its list afterwards), in the order of that point
    {
      if (qtail + nnew + nmoved - qhead > w.ring_mask + 1) {
        // compact the ring (drop the entries of facets that were deleted while pending)
        HULL_COVER(4)
        int nt = qhead;
        for (int b = qhead; b < qtail; b += 32) {
          const int pos = b + lane;
          int f = -1; bool ok = false;
          if (pos < qtail) { f = w.ring[pos & w.ring_mask]; ok = (w.mt[f].qpos == pos); }
          const unsigned m = __ballot_sync(LQRO_FULL, ok);
          __syncwarp();
          if (ok) { const int np = nt + __popc(m & lt); w.ring[np & w.ring_mask] = (hidx)f; w.mt[f].qpos = np; }
          nt += __popc(m);
          __syncwarp();
        }
        qtail = nt;
        if (qtail + nnew + nmoved - qhead > w.ring_mask + 1) { failc = 2; break; }
      }
      for (int q0 = 0; q0 < nnew; q0 += 32) {
        const int q = q0 + lane;
        int f = 0; bool ne = false;
        if (q < nnew) { f = w.newf[q]; ne = w.ot[f].ocount > 0; }
        const unsigned m = __ballot_sync(LQRO_FULL, ne);
        if (ne) { const int pos = qtail + __popc(m & lt); w.ring[pos & w.ring_mask] = (hidx)f; w.mt[f].qpos = pos; }
        qtail += __popc(m);
      }
      for (int j = lane; j < nmoved; j += 32) {
        const int f = w.vis[nvis + j];
        const int pos = qtail + j;
        w.ring[pos & w.ring_mask] = (hidx)f;
        HMeta M; M.seq = seqctr + 1 + j; M.qpos = pos;
        w.mt[f] = M;
      }
      qtail += nmoved; seqctr += nmoved;
      // qh_deletevisible: the visible facets leave the hull, their ids go to the free list (two warps per hull: done above)
      if (!COOP) {
        for (int vi = lane; vi < nvis; vi += 32) {
          const int vf = w.vis[vi];
          w.tp[vf].born = -1;
          w.mt[vf].qpos = -1;
          w.freel[nfree + vi] = (hidx)vf;
        }
        nfree += nvis;
      }
      __syncwarp();
    }
    HULL_TIC(9)
  }
  if (COOP) {  // leave the helper at its request barrier with "hull finished"
    if (hstate == 1) { hull_pair_bar(mail->bar_id); hstate = 2; }
    if (hstate == 2) hull_pair_bar(mail->bar_id);
    if (lane == 0) mail->cmd = 0;
    hull_pair_bar(mail->bar_id);
  }
  // ---- results
  int nf = 0;
  if (!failc) {
    for (int f0 = 0; f0 < pool_used; f0 += 32) {
      const int f = f0 + lane;
      const bool alive = f < pool_used && w.tp[f].born >= 0;
      nf += __popc(__ballot_sync(LQRO_FULL, alive));
    }
    if (want_order) {
      // position in qhull's facet list = number of alive facets with a smaller sequence key
      for (int f = lane; f < pool_used; f += 32) {
        if (w.tp[f].born < 0) continue;
        const int s = w.mt[f].seq;
        int r = 0;
        for (int g = 0; g < pool_used; ++g) r += (w.tp[g].born >= 0 && w.mt[g].seq < s) ? 1 : 0;
        w.order[r] = (hidx)f;
      }
    }
  }
  if (__any_sync(LQRO_FULL, degen)) degen = 1;
  if (lane == 0) {
    if (degen) c.degenerate = 1;
    c.fail = failc; c.nfacets_out = nf; c.ninserted = ins; c.pool_used = pool_used;
#if defined(LQRO_HULL_PROF)
    for (int q = 0; q < 12; ++q) c.prof[q] = hprof[q];
#endif
  }
  __syncwarp();
}

__device__ __forceinline__ void hull_build(HullWs& w, HullCtl& c, int n, bool want_order) { hull_build_t<false>(w, c, nullptr, n, want_order); }

// Nearest facet to v_rel in qhull's facet order (MAIN:825-837, with normal = facet 0's when it wins):
// distance_f = |n_f . (v_rel - p_f)|, p_f = first listed vertex of f from the UN-ROUNDED cloud (ux,uy,uz);
// ties go to the facet that comes first in qhull's list (smaller sequence key).  Returns the facet id.
__device__ __forceinline__ int hull_nearest_facet(const HullWs& w, const HullCtl& c, const double* ux, const double* uy, const double* uz,
                                                  double vr0, double vr1, double vr2, double* distance, double* normal) {
  const int lane = lane_id();
  double bd = 1.7976931348623157e308;
  int bs = 0x7fffffff, bf = 0;
  const int np = c.pool_used;
  for (int f = lane; f < np; f += 32) {
    const HTopo T = w.tp[f];
    if (T.born < 0) continue;
    const HPlane F = w.pl[f];
    const int p = w.vpoint[T.v[0]];
    const double d = fabs(F.n0 * (vr0 - ux[p]) + F.n1 * (vr1 - uy[p]) + F.n2 * (vr2 - uz[p]));
    const int s = w.mt[f].seq;
    if (d < bd || (d == bd && s < bs)) { bd = d; bs = s; bf = f; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(LQRO_FULL, bd, off);
    const int os = __shfl_xor_sync(LQRO_FULL, bs, off);
    const int of = __shfl_xor_sync(LQRO_FULL, bf, off);
    if (ov < bd || (ov == bd && os < bs)) { bd = ov; bs = os; bf = of; }
  }
  *distance = bd;
  const HPlane F = w.pl[bf];
  normal[0] = F.n0; normal[1] = F.n1; normal[2] = F.n2;
  return bf;
}
// position of facet f in qhull's facet list (all lanes)
__device__ __forceinline__ int hull_list_position(const HullWs& w, const HullCtl& c, int f) {
  const int lane = lane_id();
  const int s = w.mt[f].seq;
  int r = 0;
  for (int g = lane; g < c.pool_used; g += 32) r += (w.tp[g].born >= 0 && w.mt[g].seq < s) ? 1 : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(LQRO_FULL, r, off);
  return r;
}
