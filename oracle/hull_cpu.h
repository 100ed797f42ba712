// oracle/hull_cpu.h -- TEST INFRASTRUCTURE ONLY.
// CPU convex hull used (a) as the in-process stand-in for the reference's qconvex.exe spawn
// (MAIN:769-770), (b) as the checker for the CUDA hull, (c) as the hull of the CPU baseline.
//
// The reference obtains its hull from qhull 2012.1 (QuadrotorHoverController/qconvex.exe, a Windows
// PE that cannot run here; qhull is NOT vendored in /root/reference).  Two things the reference
// consumes depend on qhull's internal processing order, not only on the geometry:
//   * the ORDER of facets in Planes.txt / facetVertices.txt (ties in the nearest-facet scan,
//     MAIN:829-837, and the facet-0 stale-normal defect), and
//   * the FIRST vertex listed for each facet (MAIN:799-802: the point the plane distance is
//     measured from, taken from the un-rounded cloud while the normal comes from the 6-digit
//     cloud, so different vertices of one facet give distances that differ by ~1e-5 relative).
// lqro_hull_qorder() therefore re-states qhull's published Quickhull procedure (Barber, Dobkin,
// Huhdanpaa 1996; qhull libqhull: qh_maxmin, qh_maxsimplex, qh_createsimplex, qh_partitionall,
// qh_nextfurthest, qh_findhorizon, qh_makenew_simplicial, qh_findbest, qh_partitionpoint) for
// simplicial, non-degenerate 3-d input, including its vertex numbering (a facet lists its vertices
// by decreasing creation id, so the first vertex is the apex the facet was built from) and its
// facet-list order.  It is pinned against the reference's own fixture (pointList.txt ->
// Planes.txt + facetVertices.txt: same facets, same order, same vertex order) and against
// scipy.spatial.ConvexHull (qhull_r 8.0.2) on generated clouds -- see tests/test_oracle_hull.py.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lqro_hull_result {
  int nfacets;
  int nvertices;      // hull vertices
  int ninserted;      // points processed (>= nvertices)
  int degenerate;     // count of decisions that fell inside the round-off band (0 for generic input)
  double* planes;     // nfacets x 4: outward unit normal, offset (n.x + off = 0 on the facet)
  int* verts;         // nfacets x 3: point indices, qhull 'Fv' order (apex first)
  int* insert_order;  // ninserted point indices in processing order
} lqro_hull_result;

// returns 0 on success, <0 on failure (-1: fewer than 4 points / flat input, -2: internal limit)
int lqro_hull_qorder(const double* pts, int n, lqro_hull_result* out);
void lqro_hull_free(lqro_hull_result* r);

// printf("%g")/strtod round trip of MAIN:765-767 for one coordinate: the double nearest to the
// 6-significant-digit decimal rendering of x.
double lqro_round6(double x);

#ifdef __cplusplus
}
#endif
