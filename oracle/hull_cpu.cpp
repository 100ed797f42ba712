// oracle/hull_cpu.cpp -- TEST INFRASTRUCTURE ONLY.  See hull_cpu.h for what this is and why.
//
// Re-statement of qhull's Quickhull for simplicial, non-degenerate 3-d input.  qhull is a
// third-party dependency of the reference (shipped only as qconvex.exe, qhull 2012.1); the
// routine names cited below are qhull's (libqhull/{geom,geom2,poly,poly2,libqhull}.c), whose
// algorithm is re-stated from its published description, not copied.  Merging ('C-0'), coplanar
// points and narrow-hull handling are not re-stated: inputs that would exercise them raise the
// `degenerate` counter instead.
#include "hull_cpu.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Facet {
  int v[3];   // vertex ids, decreasing (qhull keeps facet->vertices sorted by decreasing id)
  int nb[3];  // neighbour opposite v[i]
  bool toporient, visible, isnew, alive;
  double n[3], off;
  int prev, next;
  std::vector<int> outside;  // point indices, furthest last (qh_partitionall / qh_partitionpoint)
  double furthestdist;
  unsigned visitid;
  int replace;
};

struct Hull {
  const double* P;
  int n;
  std::vector<Facet> F;     // pool, index == creation order; F[0] is the list sentinel (qh facet_tail)
  std::vector<int> vpoint;  // vertex id -> point index
  int facet_list, facet_next, newfacet_list, visible_list;
  unsigned visit_id;
  double DISTround, MINvisible, MINoutside, MAXcoplanar, max_outside, NEARzero[3], MINdenom;
  double interior[3];
  int degenerate;
  bool findbestnew, findbest_notsharp;
  std::vector<int> insert_order;

  const double* pt(int i) const { return P + 3 * (size_t)i; }
  double DISToutside() const { double a = 2 * MINoutside; return a > max_outside ? a : max_outside; }  // qh_DISToutside (MERGING)
  double SEARCHdist() const { double a = MINvisible > MAXcoplanar ? MINvisible : MAXcoplanar; return max_outside + 2 * DISTround + a; }

  // qh_distplane, 3-d
  double dist(const double* p, const Facet& f) const { return f.off + p[0] * f.n[0] + p[1] * f.n[1] + p[2] * f.n[2]; }

  // qh_appendfacet / qh_removefacet
  void append(int f) {
    const int tail = 0;
    if (tail == newfacet_list) newfacet_list = f;
    if (tail == facet_next) facet_next = f;
    F[f].prev = F[tail].prev;
    F[f].next = tail;
    if (F[tail].prev >= 0) F[F[tail].prev].next = f; else facet_list = f;
    F[tail].prev = f;
  }
  void remove(int f) {
    int nx = F[f].next, pv = F[f].prev;
    if (f == newfacet_list) newfacet_list = nx;
    if (f == facet_next) facet_next = nx;
    if (f == visible_list) visible_list = nx;
    if (pv >= 0) { F[pv].next = nx; F[nx].prev = pv; }
    else { facet_list = nx; F[nx].prev = -1; }
  }
  int newfacet() {
    Facet f;
    f.toporient = f.visible = f.isnew = false; f.alive = true;
    f.prev = f.next = -1; f.furthestdist = 0; f.visitid = 0; f.replace = -1;
    f.v[0] = f.v[1] = f.v[2] = -1; f.nb[0] = f.nb[1] = f.nb[2] = -1;
    f.n[0] = f.n[1] = f.n[2] = f.off = 0;
    F.push_back(f);
    return (int)F.size() - 1;
  }

  // qh_setfacetplane -> qh_sethyperplane_det (dim 3) + qh_normalize2
  void setplane(int fi) {
    Facet& f = F[fi];
    const double* p0 = pt(vpoint[f.v[0]]);
    const double* p1 = pt(vpoint[f.v[1]]);
    const double* p2 = pt(vpoint[f.v[2]]);
    double dX1 = p1[0] - p0[0], dY1 = p1[1] - p0[1], dZ1 = p1[2] - p0[2];
    double dX2 = p2[0] - p0[0], dY2 = p2[1] - p0[1], dZ2 = p2[2] - p0[2];
    double n0 = dY2 * dZ1 - dZ2 * dY1;
    double n1 = dX1 * dZ2 - dZ1 * dX2;
    double n2 = dX2 * dY1 - dY2 * dX1;
    double norm = std::sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    if (!(norm > MINdenom)) ++degenerate;
    if (!f.toporient) norm = -norm;
    f.n[0] = n0 / norm; f.n[1] = n1 / norm; f.n[2] = n2 / norm;
    f.off = -(p0[0] * f.n[0] + p0[1] * f.n[1] + p0[2] * f.n[2]);
  }
};

double det2(double a1, double a2, double b1, double b2) { return a1 * b2 - a2 * b1; }
double det3(double a1, double a2, double a3, double b1, double b2, double b3, double c1, double c2, double c3) {
  return a1 * det2(b2, b3, c2, c3) - b1 * det2(a2, a3, c2, c3) + c1 * det2(a2, a3, b2, b3);
}

// qh_detsimplex: determinant of (simplex[i] - apex) using the first k coordinates, k = simplex size
double detsimplex(const Hull& h, int apex, const int* simplex, int k, bool* nearzero) {
  const double* a = h.pt(apex);
  double r[3][3];
  for (int i = 0; i < k; ++i) {
    const double* p = h.pt(simplex[i]);
    for (int c = 0; c < k; ++c) r[i][c] = p[c] - a[c];
  }
  double det;
  *nearzero = false;
  if (k == 2) {
    det = det2(r[0][0], r[0][1], r[1][0], r[1][1]);
    if (std::fabs(det) < h.NEARzero[1]) *nearzero = true;
  } else {
    det = det3(r[0][0], r[0][1], r[0][2], r[1][0], r[1][1], r[1][2], r[2][0], r[2][1], r[2][2]);
    if (std::fabs(det) < h.NEARzero[2]) *nearzero = true;
  }
  return det;
}

// qh_findbesthorizon (ischeckmax = false): flood over neighbours that are within SEARCHdist of the best
int findbesthorizon(Hull& h, const double* p, int startfacet, double* bestdist) {
  int bestfacet = startfacet;
  unsigned visitid = ++h.visit_id;
  double searchdist = h.SEARCHdist();
  double minsearch = *bestdist - searchdist;
  std::vector<int> coplanarset;
  int facet = startfacet;
  int nextfacet = -1;
  h.F[startfacet].visitid = visitid;
  while (true) {
    Facet& f = h.F[facet];
    for (int k = 0; k < 3; ++k) {
      int nbi = f.nb[k];
      if (nbi < 0) continue;
      Facet& nb = h.F[nbi];
      if (nb.visitid == visitid) continue;
      nb.visitid = visitid;
      double d = h.dist(p, nb);
      if (d > *bestdist) {
        bestfacet = nbi;
        if (d > *bestdist + searchdist) coplanarset.clear();
        *bestdist = d;
        minsearch = d - searchdist;
      } else if (d < minsearch) {
        continue;
      }
      if (nextfacet >= 0) coplanarset.push_back(nextfacet);
      nextfacet = nbi;
    }
    facet = nextfacet;
    if (facet >= 0) nextfacet = -1;
    else if (coplanarset.empty()) break;
    else { facet = coplanarset.back(); coplanarset.pop_back(); }
  }
  return bestfacet;
}

// qh_findbestnew: exhaustive over the new facets, cyclically from startfacet
int findbestnew(Hull& h, const double* p, int startfacet, double* dist, bool* isoutside) {
  double bestdist = -DBL_MAX / 2;
  int bestfacet = -1;
  double distoutside = h.DISToutside();
  *isoutside = true;
  ++h.visit_id;
  int facet = startfacet;
  for (int i = 0; i < 2; ++i, facet = h.newfacet_list) {
    for (; facet != 0; facet = h.F[facet].next) {
      if (facet == startfacet && i) break;
      double d = h.dist(p, h.F[facet]);
      if (d > bestdist) {
        bestfacet = facet;
        if (d >= distoutside) { *dist = d; return bestfacet; }
        bestdist = d;
      }
    }
  }
  bestfacet = findbesthorizon(h, p, bestfacet >= 0 ? bestfacet : startfacet, &bestdist);
  *dist = bestdist;
  if (bestdist < h.MINoutside) *isoutside = false;
  return bestfacet;
}

// qh_sharpnewfacets: do the new facets' normals leave one orthant?
bool sharpnewfacets(const Hull& h) {
  bool quadrant[3] = {false, false, false};
  for (int f = h.newfacet_list; f != 0; f = h.F[f].next) {
    if (f == h.newfacet_list) {
      for (int k = 0; k < 3; ++k) quadrant[k] = h.F[f].n[k] > 0;
    } else {
      for (int k = 0; k < 3; ++k)
        if (quadrant[k] != (h.F[f].n[k] > 0)) return true;
    }
  }
  return false;
}

// qh_findbest(point, startfacet, !bestoutside, isnewfacets, ...): directed walk over the new facets
int findbest_new(Hull& h, const double* p, int startfacet, double* dist, bool* isoutside) {
  double bestdist = -DBL_MAX / 2;
  int bestfacet = -1;
  unsigned visitid = ++h.visit_id;
  bool testhorizon = true;
  *isoutside = true;
  double d = h.dist(p, h.F[startfacet]);
  if (d >= h.MINoutside) { *dist = d; return startfacet; }
  bestdist = d;
  bestfacet = startfacet;
  h.F[startfacet].visitid = visitid;
  int facet = startfacet;
  while (facet >= 0) {
    Facet& f = h.F[facet];
    int nextf = -1;
    for (int k = 0; k < 3; ++k) {
      int nbi = f.nb[k];
      Facet& nb = h.F[nbi];
      if (!nb.isnew) continue;
      if (nb.visitid == visitid) continue;
      nb.visitid = visitid;
      d = h.dist(p, nb);
      if (d > bestdist) {
        if (d >= h.MINoutside) { *dist = d; return nbi; }
        bestfacet = nbi;
        bestdist = d;
        nextf = nbi;
        break;  // switch to neighbour
      }
    }
    facet = nextf;
  }
  if (!h.findbest_notsharp && bestdist < -h.DISTround) {
    if (sharpnewfacets(h)) {
      bestfacet = findbestnew(h, p, bestfacet, &bestdist, isoutside);
      testhorizon = false;
      h.findbestnew = true;
      *dist = bestdist;
      return bestfacet;  // findbestnew already applied the horizon test and isoutside rule
    } else {
      h.findbest_notsharp = true;
    }
  }
  if (testhorizon) bestfacet = findbesthorizon(h, p, bestfacet, &bestdist);
  *dist = bestdist;
  if (bestdist < h.MINoutside) *isoutside = false;
  return bestfacet;
}

// qh_partitionpoint
void partitionpoint(Hull& h, int pi, int startfacet) {
  const double* p = h.pt(pi);
  double bestdist;
  bool isoutside;
  int best;
  if (h.findbestnew) best = findbestnew(h, p, startfacet, &bestdist, &isoutside);
  else best = findbest_new(h, p, startfacet, &bestdist, &isoutside);
  if (std::fabs(bestdist) < 4 * h.MINoutside) ++h.degenerate;
  if (isoutside) {
    Facet& bf = h.F[best];
    if (bf.outside.empty()) {
      bf.outside.push_back(pi);
      if (!bf.isnew) { h.remove(best); h.append(best); }
      bf.furthestdist = bestdist;
    } else if (bf.furthestdist < bestdist) {
      bf.outside.push_back(pi);
      bf.furthestdist = bestdist;
    } else {
      bf.outside.insert(bf.outside.end() - 1, pi);  // qh_setappend2ndlast
    }
  }
  // else: coplanar / inside -> dropped (no 'Qc', no 'Qi')
}

}  // namespace

extern "C" double lqro_round6(double x) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.6g", x);
  return strtod(buf, 0);
}

extern "C" void lqro_hull_free(lqro_hull_result* r) {
  if (!r) return;
  free(r->planes); free(r->verts); free(r->insert_order);
  r->planes = 0; r->verts = 0; r->insert_order = 0;
}

extern "C" int lqro_hull_qorder(const double* pts, int n, lqro_hull_result* out) {
  memset(out, 0, sizeof *out);
  if (n < 4) return -1;
  Hull h;
  h.P = pts; h.n = n;
  h.visit_id = 0; h.degenerate = 0; h.findbestnew = false; h.findbest_notsharp = false;
  h.max_outside = 0;
  h.F.reserve(64 + 8 * (size_t)n);
  h.newfacet();  // sentinel (qh facet_tail)
  h.F[0].alive = false;
  h.facet_list = 0; h.facet_next = 0; h.newfacet_list = -1; h.visible_list = -1;

  // ---- qh_maxmin: extreme points per coordinate, round-off constants (qh_detroundoff, qh_distround)
  int maxpoints[6];
  double MAXabs = 0, MAXsum = 0, MAXwidth = 0;
  for (int k = 0; k < 3; ++k) {
    int mn = 0, mx = 0;
    for (int i = 0; i < n; ++i) {
      if (h.pt(mx)[k] < h.pt(i)[k]) mx = i;
      else if (h.pt(mn)[k] > h.pt(i)[k]) mn = i;
    }
    double maxcoord = h.pt(mx)[k] > -h.pt(mn)[k] ? h.pt(mx)[k] : -h.pt(mn)[k];
    double w = h.pt(mx)[k] - h.pt(mn)[k];
    if (w > MAXwidth) MAXwidth = w;
    if (maxcoord > MAXabs) MAXabs = maxcoord;
    MAXsum += maxcoord;
    maxpoints[2 * k] = mx; maxpoints[2 * k + 1] = mn;
    h.NEARzero[k] = 80 * MAXsum * DBL_EPSILON;
  }
  {
    double maxdistsum = std::sqrt(3.0) * MAXabs;
    if (MAXsum < maxdistsum) maxdistsum = MAXsum;
    h.DISTround = DBL_EPSILON * (3 * maxdistsum * 1.01 + MAXabs);
    h.MINvisible = 2 * h.DISTround;   // premerge_centrum ('C-0') + 2*DISTround
    h.MAXcoplanar = h.MINvisible;
    h.MINoutside = 2 * h.MINvisible;
    h.MINdenom = (MAXabs > 1e-300 ? MAXabs : 1e-300) * DBL_MIN / DBL_EPSILON;  // never binds for sane input
  }

  // ---- qh_maxsimplex
  int simplex[4]; int ns = 0;
  {
    double maxc = -DBL_MAX, minc = DBL_MAX; int maxx = -1, minx = -1;
    for (int i = 0; i < 6; ++i) {
      const double* p = h.pt(maxpoints[i]);
      if (maxc < p[0]) { maxc = p[0]; maxx = maxpoints[i]; }
      if (minc > p[0]) { minc = p[0]; minx = maxpoints[i]; }
    }
    simplex[ns++] = minx;
    if (maxx != minx) simplex[ns++] = maxx;
    if (ns < 2) return -1;
  }
  for (int k = ns; k < 4; ++k) {
    int maxpoint = -1; double maxdet = -DBL_MAX; bool maxnearzero = false;
    for (int i = 0; i < 6; ++i) {
      int p = maxpoints[i];
      bool in = false;
      for (int s = 0; s < k; ++s) if (simplex[s] == p) in = true;
      if (in) continue;
      bool nz; double det = std::fabs(detsimplex(h, p, simplex, k, &nz));
      if (det > maxdet) { maxdet = det; maxpoint = p; maxnearzero = nz; }
    }
    if (maxpoint < 0 || maxnearzero) {
      for (int p = 0; p < n; ++p) {
        bool in = false;
        for (int s = 0; s < k; ++s) if (simplex[s] == p) in = true;
        if (in) continue;
        bool nz; double det = std::fabs(detsimplex(h, p, simplex, k, &nz));
        if (det > maxdet) { maxdet = det; maxpoint = p; maxnearzero = nz; }
      }
    }
    if (maxpoint < 0 || maxnearzero) return -1;
    simplex[k] = maxpoint;
  }

  // ---- qh_initialvertices + qh_createsimplex + qh_initialhull
  for (int i = 0; i < 4; ++i) h.vpoint.push_back(simplex[i]);  // vertex id i <-> simplex[i]
  int verts_desc[4] = {3, 2, 1, 0};
  bool toporient = true;
  int fid[4];
  for (int i = 0; i < 4; ++i) {
    int f = h.newfacet(); fid[i] = f;
    int c = 0;
    for (int j = 0; j < 4; ++j) if (j != i) h.F[f].v[c++] = verts_desc[j];
    h.F[f].toporient = toporient;
    h.F[f].isnew = true;
    h.append(f);
    toporient = !toporient;
  }
  for (int i = 0; i < 4; ++i) {
    int c = 0;
    for (int j = 0; j < 4; ++j) if (j != i) h.F[fid[i]].nb[c++] = fid[j];
  }
  h.facet_next = h.facet_list;
  for (int k = 0; k < 3; ++k) {
    double s = 0;
    for (int j = 0; j < 4; ++j) s += h.pt(h.vpoint[verts_desc[j]])[k];
    h.interior[k] = s / 4;
  }
  h.setplane(fid[0]);
  if (h.dist(h.interior, h.F[fid[0]]) > 0)
    for (int i = 0; i < 4; ++i) h.F[fid[i]].toporient = !h.F[fid[i]].toporient;
  for (int i = 0; i < 4; ++i) h.setplane(fid[i]);
  for (int i = 0; i < 4; ++i)
    if (!(h.dist(h.interior, h.F[fid[i]]) < -h.DISTround)) ++h.degenerate;
  for (int i = 0; i < 4; ++i) h.F[fid[i]].isnew = false;  // qh_resetlists

  // ---- qh_partitionall
  {
    std::vector<int> pointset;
    pointset.reserve(n);
    for (int i = 0; i < n; ++i) {
      bool isv = false;
      for (int s = 0; s < 4; ++s) if (simplex[s] == i) isv = true;
      if (!isv) pointset.push_back(i);
    }
    double distoutside = h.DISToutside();
    for (int f = h.facet_list; f != 0; f = h.F[f].next) {
      Facet& fc = h.F[f];
      int bestpoint = -1; double bestdist = -DBL_MAX;
      size_t keep = 0;
      for (size_t q = 0; q < pointset.size(); ++q) {
        int pi = pointset[q];
        double d = h.dist(h.pt(pi), fc);
        if (std::fabs(d) < 2 * distoutside) ++h.degenerate;
        if (d < distoutside) pointset[keep++] = pi;
        else {
          if (bestpoint < 0) { bestpoint = pi; bestdist = d; }
          else if (d > bestdist) { fc.outside.push_back(bestpoint); bestpoint = pi; bestdist = d; }
          else fc.outside.push_back(pi);
        }
      }
      if (bestpoint >= 0) { fc.outside.push_back(bestpoint); fc.furthestdist = bestdist; }
      pointset.resize(keep);
    }
  }

  // ---- qh_furthestnext (called once from qh_initbuild): the initial facet whose furthest point is
  //      furthest of all moves to the front of the facet list
  {
    int bestfacet = -1; double bestdist = -DBL_MAX;
    for (int f = h.facet_list; f != 0; f = h.F[f].next)
      if (!h.F[f].outside.empty() && h.F[f].furthestdist > bestdist) { bestfacet = f; bestdist = h.F[f].furthestdist; }
    if (bestfacet >= 0 && bestfacet != h.facet_list) {
      h.remove(bestfacet);
      int head = h.facet_list;  // qh_prependfacet(bestfacet, &qh facet_next), facet_next == facet_list here
      h.F[bestfacet].prev = -1; h.F[bestfacet].next = head; h.F[head].prev = bestfacet;
      h.facet_list = bestfacet;
    }
    h.facet_next = h.facet_list;
  }

  // ---- qh_buildhull: qh_nextfurthest / qh_addpoint
  std::vector<int> visible, newf, horizon_of_new;
  const size_t max_pool = 64 + 16 * (size_t)n;
  while (true) {
    // qh_nextfurthest
    int facet = -1;
    while ((facet = h.facet_next) != 0) {
      if (h.F[facet].outside.empty()) { h.facet_next = h.F[facet].next; continue; }
      break;
    }
    if (facet == 0) break;
    int furthest = h.F[facet].outside.back();
    h.F[facet].outside.pop_back();
    h.insert_order.push_back(furthest);
    const double* ap = h.pt(furthest);

    // qh_findhorizon
    visible.clear();
    h.remove(facet); h.append(facet);
    h.visible_list = facet;
    h.F[facet].visible = true; h.F[facet].replace = -1;
    visible.push_back(facet);
    unsigned vid = ++h.visit_id;
    for (size_t vi = 0; vi < visible.size(); ++vi) {
      int vf = visible[vi];
      h.F[vf].visitid = vid;
      for (int k = 0; k < 3; ++k) {
        int nbi = h.F[vf].nb[k];
        Facet& nb = h.F[nbi];
        if (nb.visitid == vid) continue;
        nb.visitid = vid;
        double d = h.dist(ap, nb);
        if (d > h.MINvisible) {
          h.remove(nbi); h.append(nbi);
          nb.visible = true; nb.replace = -1;
          visible.push_back(nbi);
        } else if (d > -h.MAXcoplanar) {
          ++h.degenerate;  // coplanar horizon: qhull would merge
        }
      }
    }

    // qh_makenewfacets / qh_makenew_simplicial
    h.newfacet_list = 0;
    int apex = (int)h.vpoint.size();
    h.vpoint.push_back(furthest);
    newf.clear();
    for (size_t vi = 0; vi < visible.size(); ++vi) {
      int vf = visible[vi];
      int last_new = -1;
      for (int k = 0; k < 3; ++k) {
        int nbi = h.F[vf].nb[k];
        if (h.F[nbi].visible) continue;
        // horizon facet nbi, ridge = its vertices minus the one opposite vf (qh_facetintersect)
        int horizonskip = -1;
        for (int q = 0; q < 3; ++q) if (h.F[nbi].nb[q] == vf) { horizonskip = q; break; }
        if (horizonskip < 0) return -2;
        if (h.F.size() >= max_pool) return -2;
        int nf = h.newfacet();
        Facet& hz = h.F[nbi];
        Facet& f = h.F[nf];
        f.v[0] = apex;
        int c = 1;
        for (int q = 0; q < 3; ++q) if (q != horizonskip) f.v[c++] = hz.v[q];
        f.toporient = hz.toporient ? (horizonskip & 1) : ((horizonskip & 1) ^ 1);
        f.nb[0] = nbi;
        f.isnew = true;
        h.append(nf);
        hz.nb[horizonskip] = nf;
        newf.push_back(nf);
        last_new = nf;
      }
      if (last_new >= 0) h.F[vf].replace = last_new;
    }
    // qh_makenewplanes
    for (size_t i = 0; i < newf.size(); ++i) {
      h.setplane(newf[i]);
      if (!(h.dist(h.interior, h.F[newf[i]]) < 0)) ++h.degenerate;
    }
    // qh_matchnewfacets: new facets share the ridges {apex, w}; each ridge vertex w lies on two new facets
    {
      // for vertex w: facets where w == v[1] (its neighbour opposite v[2] shares {apex, v[1]}) etc.
      // neighbour opposite v[1] shares ridge {apex, v[2]}; neighbour opposite v[2] shares ridge {apex, v[1]}
      size_t m = newf.size();
      for (size_t a = 0; a < m; ++a) {
        Facet& fa = h.F[newf[a]];
        for (int q = 1; q <= 2; ++q) {
          int w = fa.v[3 - q];  // shared vertex for the neighbour opposite v[q]
          int found = -1;
          for (size_t b = 0; b < m; ++b) {
            if (b == a) continue;
            const Facet& fb = h.F[newf[b]];
            if (fb.v[1] == w || fb.v[2] == w) { found = newf[b]; break; }
          }
          if (found < 0) return -2;
          fa.nb[q] = found;
        }
      }
    }
    // qh_partitionvisible
    h.findbestnew = false; h.findbest_notsharp = false;
    for (size_t vi = 0; vi < visible.size(); ++vi) {
      Facet& vf = h.F[visible[vi]];
      if (vf.outside.empty()) continue;
      int start = vf.replace >= 0 ? vf.replace : h.newfacet_list;
      std::vector<int> pts_to_do;
      pts_to_do.swap(vf.outside);
      for (size_t q = 0; q < pts_to_do.size(); ++q) partitionpoint(h, pts_to_do[q], start);
    }
    h.findbestnew = false; h.findbest_notsharp = false;
    // qh_deletevisible + qh_resetlists
    for (size_t vi = 0; vi < visible.size(); ++vi) {
      h.remove(visible[vi]);
      h.F[visible[vi]].alive = false;
    }
    for (size_t i = 0; i < newf.size(); ++i) h.F[newf[i]].isnew = false;
    h.newfacet_list = -1; h.visible_list = -1;
  }

  // ---- output in facet-list order ('n' and 'Fv' both walk qh facet_list)
  int nf = 0;
  for (int f = h.facet_list; f != 0; f = h.F[f].next) ++nf;
  out->nfacets = nf;
  out->planes = (double*)malloc(sizeof(double) * 4 * (size_t)nf);
  out->verts = (int*)malloc(sizeof(int) * 3 * (size_t)nf);
  out->ninserted = (int)h.insert_order.size();
  out->insert_order = (int*)malloc(sizeof(int) * (h.insert_order.size() + 1));
  for (size_t i = 0; i < h.insert_order.size(); ++i) out->insert_order[i] = h.insert_order[i];
  std::vector<char> isvert(n, 0);
  int c = 0;
  for (int f = h.facet_list; f != 0; f = h.F[f].next, ++c) {
    const Facet& fc = h.F[f];
    for (int k = 0; k < 3; ++k) {
      out->planes[4 * c + k] = fc.n[k];
      out->verts[3 * c + k] = h.vpoint[fc.v[k]];
      isvert[h.vpoint[fc.v[k]]] = 1;
    }
    out->planes[4 * c + 3] = fc.off;
  }
  int nv = 0;
  for (int i = 0; i < n; ++i) nv += isvert[i];
  out->nvertices = nv;
  out->degenerate = h.degenerate;
  return 0;
}
