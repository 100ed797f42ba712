// oracle/lqro_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  See lqro_oracle.h.
// Compile with -ffp-contract=off: every product and sum below is rounded separately, as in the
// reference built without FMA contraction.
#include "lqro_oracle.h"
#include "hull_cpu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <limits>
#include <thread>
#include <utility>
#include <vector>

namespace {

const int XD = LQRO_O_XDIM;

// ---------------------------------------------------------------------------------------------
// Small dense helpers in the reference's operation order.
// MAT:218-231: C(i,j) = sum_k A(i,k)*B(k,j), accumulated left to right from 0.
void matmul(const double* A, const double* B, double* C, int n, int m, int p) {  // (n x m) * (m x p)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < p; ++j) {
      double t = 0;
      for (int k = 0; k < m; ++k) t += A[i * m + k] * B[k * p + j];
      C[i * p + j] = t;
    }
}
void transpose(const double* A, double* At, int n, int m) {
  for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) At[j * n + i] = A[i * m + j];
}

// MAT:603-671 `operator!`: Gauss-Jordan with full pivoting, row/column permutation vectors.
void inverse(const double* q, double* out, int n) {
  std::vector<double> m(q, q + n * n), inv(n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[i * n + i] = 1.0;
  std::vector<int> rp(n), cp(n);
  for (int i = 0; i < n; ++i) { rp[i] = i; cp[i] = i; }
  for (int k = 0; k < n; ++k) {
    double maximum = 0; int mr = k, mc = k;
    for (int i = k; i < n; ++i)
      for (int j = k; j < n; ++j) {
        double a = std::fabs(m[rp[i] * n + cp[j]]);
        if (a > maximum) { maximum = a; mr = i; mc = j; }
      }
    std::swap(rp[k], rp[mr]);
    std::swap(cp[k], cp[mc]);
    for (int i = k + 1; i < n; ++i) {
      double factor = m[rp[i] * n + cp[k]] / m[rp[k] * n + cp[k]];
      for (int j = k + 1; j < n; ++j) m[rp[i] * n + cp[j]] -= factor * m[rp[k] * n + cp[j]];
      for (int j = 0; j < k; ++j) inv[rp[i] * n + rp[j]] -= factor * inv[rp[k] * n + rp[j]];
      inv[rp[i] * n + rp[k]] = -factor;
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double quotient = m[rp[k] * n + cp[k]];
    for (int j = 0; j < n; ++j) inv[rp[k] * n + j] /= quotient;
    for (int i = 0; i < k; ++i) {
      double factor = m[rp[i] * n + cp[k]];
      for (int j = 0; j < n; ++j) inv[rp[i] * n + j] -= factor * inv[rp[k] * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[cp[i] * n + j] = inv[rp[i] * n + j];
}

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }  // GJK.h:36-38

// ---------------------------------------------------------------------------------------------
// GJK (Cameron's v2.4 distance routine) for a point against a cloud.
// Subsets of the <=4 simplex points are bit masks; the element / predecessor / successor tables of
// GJK.c:88-159 are what these helpers enumerate (elements ascending).
inline int card(int s) { return __builtin_popcount(s); }
inline int elt(int s, int i) {  // i-th set bit
  for (int b = 0; b < 4; ++b) if (s & (1 << b)) { if (i == 0) return b; --i; }
  return 0;
}
inline int non_elt(int s, int i) {  // i-th clear bit
  for (int b = 0; b < 4; ++b) if (!(s & (1 << b))) { if (i == 0) return b; --i; }
  return 0;
}

struct Gjk {
  int npts;
  int simplex2[4];
  double lambdas[4];
  double c1[4][3], c2[4][3];
  double del[16][4];   // delta_values[s][elt]
  double dsum[16];     // delta[s]
  double prod[4][4];
  int st[6];           // g_test, simplices, backups, dot_products, support_dp, other_ops
};

void compute_subterms(Gjk& g) {  // GJK.c:527-581
  int size = g.npts;
  double c[4][3];
  for (int i = 0; i < size; ++i) for (int j = 0; j < 3; ++j) c[i][j] = g.c1[i][j] - g.c2[i][j];
  for (int i = 0; i < size; ++i)
    for (int j = i; j < size; ++j) { g.prod[i][j] = g.prod[j][i] = dot3(c[i], c[j]); ++g.st[3]; }
  for (int s = 1; s < (1 << size); ++s) {
    int cs = card(s);
    if (cs <= 1) { g.del[s][elt(s, 0)] = 1.0; continue; }
    if (cs == 2) {
      int e0 = elt(s, 0), e1 = elt(s, 1);
      g.del[s][e0] = g.prod[e1][e1] - g.prod[e1][e0];
      g.del[s][e1] = g.prod[e0][e0] - g.prod[e0][e1];
      continue;
    }
    for (int j = 0; j < cs; ++j) {
      int jelt = elt(s, j);
      int jsub = s & ~(1 << jelt);
      int e0 = elt(jsub, 0);
      double sum = 0;
      for (int i = 0; i < card(jsub); ++i) {
        int ielt = elt(jsub, i);
        sum += g.del[jsub][ielt] * (g.prod[ielt][e0] - g.prod[ielt][jelt]);
        ++g.st[5];
      }
      g.del[s][jelt] = sum;
    }
  }
}

void reset_simplex(Gjk& g, int subset) {  // GJK.c:708-736
  int cs = card(subset);
  for (int j = 0; j < cs; ++j) {
    int oldpos = elt(subset, j);
    if (oldpos != j) {
      g.simplex2[j] = g.simplex2[oldpos];
      for (int i = 0; i < 3; ++i) { g.c1[j][i] = g.c1[oldpos][i]; g.c2[j][i] = g.c2[oldpos][i]; }
    }
    g.lambdas[j] = g.del[subset][oldpos] / g.dsum[subset];
    ++g.st[5];
  }
  g.npts = cs;
}

int default_distance(Gjk& g) {  // GJK.c:593-657
  int size = g.npts;
  ++g.st[1];
  int s, ok = 0;
  for (s = 1; s < (1 << size); ++s) {
    g.dsum[s] = 0; ok = 1;
    int cs = card(s);
    for (int j = 0; ok && j < cs; ++j) {
      double d = g.del[s][elt(s, j)];
      if (d > 0) g.dsum[s] += d; else ok = 0;
    }
    for (int k = 0; ok && k < size - cs; ++k) {
      int ne = non_elt(s, k);
      if (g.del[s | (1 << ne)][ne] > 0) ok = 0;
    }
    if (ok && g.dsum[s] >= 1.0e-20) break;  // TINY, GJK.h:214
  }
  if (ok && s < (1 << size)) { reset_simplex(g, s); return 1; }
  return 0;
}

void backup_distance(Gjk& g) {  // GJK.c:663-706
  int size = g.npts, bests = 0;
  double num[16], den[16];
  ++g.st[2];
  for (int s = 1; s < (1 << size); ++s) {
    if (g.dsum[s] <= 0) continue;
    int cs = card(s), i;
    for (i = 0; i < cs; ++i) if (g.del[s][elt(s, i)] <= 0) break;
    if (i < cs) continue;
    num[s] = 0;
    for (int j = 0; j < cs; ++j)
      for (int k = 0; k < cs; ++k) {
        num[s] += (g.del[s][elt(s, j)] * g.del[s][elt(s, k)]) * g.prod[elt(s, j)][elt(s, k)];
        g.st[5] += 2;
      }
    den[s] = g.dsum[s] * g.dsum[s]; ++g.st[5];
    if (bests < 1 || num[s] * den[bests] < num[bests] * den[s]) bests = s;
  }
  reset_simplex(g, bests);
}

double gjk_point_cloud(const double* vrel, const double* pts, int n, double* w1, double* w2, int* stats) {
  Gjk g;
  memset(&g, 0, sizeof g);
  const double EPS = 1.0e-8;  // GJK.h:208
  g.npts = 1; g.simplex2[0] = 0; g.lambdas[0] = 1.0;           // GJK.c:340-347
  for (int d = 0; d < 3; ++d) { g.c1[0][d] = vrel[d]; g.c2[0][d] = pts[d]; }
  int use_default = 1, first_iteration = 1;
  long max_iterations = n;                                       // GJK.c:367, NumVertices(obj1) == 1
  double oldsqrd = 0, sqrd = 0;
  double disp[3], rdisp[3];
  while (max_iterations-- > 0) {
    if (g.npts == 1) g.lambdas[0] = 1.0;
    else {
      compute_subterms(g);
      if (use_default) use_default = default_distance(g);
      if (!use_default) backup_distance(g);
    }
    for (int d = 0; d < 3; ++d) {                                // compute_point, GJK.c:851-862
      double a = 0, b = 0;
      for (int i = 0; i < g.npts; ++i) a += g.c1[i][d] * g.lambdas[i];
      for (int i = 0; i < g.npts; ++i) b += g.c2[i][d] * g.lambdas[i];
      w1[d] = a; w2[d] = b;
    }
    g.st[5] += 6 * g.npts;
    for (int d = 0; d < 3; ++d) { disp[d] = w2[d] - w1[d]; rdisp[d] = -disp[d]; }
    sqrd = dot3(disp, disp); ++g.st[3];
    if (sqrd < EPS) break;                                       // GJK.c:419-422
    double maxv = dot3(vrel, disp);                              // support of the single point
    int minp = 0;
    double minus_minv = dot3(pts, rdisp);                        // support_simple, GJK.c:770-794
    for (int p = 1; p < n; ++p) {
      double v = dot3(pts + 3 * p, rdisp);
      if (v > minus_minv) { minus_minv = v; minp = p; }
    }
    g.st[4] += n + 2; g.st[3] += n + 2;                          // (the 1-point object is hill-climbed: 2 dots)
    ++g.st[0];
    double g_val = sqrd + maxv + minus_minv;                     // GJK.c:450
    if (g_val < 0.0) g_val = 0;
    if (g_val < EPS) break;
    if ((first_iteration || sqrd < oldsqrd) && g.npts <= 3) {    // GJK.c:472-487
      int k = g.npts;
      g.simplex2[k] = minp; g.lambdas[k] = 0;
      for (int d = 0; d < 3; ++d) { g.c1[k][d] = vrel[d]; g.c2[k][d] = pts[3 * minp + d]; }
      g.npts++;
      oldsqrd = sqrd; first_iteration = 0; use_default = 1;
      continue;
    }
    if (use_default) use_default = 0;                            // GJK.c:490-496
    else break;
  }
  if (stats) for (int i = 0; i < 6; ++i) stats[i] = g.st[i];
  if (max_iterations < 0) return 0.0;  // iteration cap exhausted: GJK.c:499 (never seen in practice)
  return sqrd;
}

// ---------------------------------------------------------------------------------------------
// FP32 3-vector and the RVO2-3D linear programs (MAIN:872-1082, V3:76-427).
struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { V3 v = {x, y, z}; return v; }
inline float dotf(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 add(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(const V3& a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 divs(const V3& a, float s) { const float inv = 1.0f / s; return mk(a.x * inv, a.y * inv, a.z * inv); }  // V3:206-211
inline V3 crossf(const V3& a, const V3& b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float absSq(const V3& a) { return dotf(a, a); }
inline V3 normalize(const V3& a) { return divs(a, std::sqrt(dotf(a, a))); }
inline float sqr(float s) { return s * s; }
const float RVO_EPS = 0.00001f;  // MAIN:841

struct PlaneF { V3 point, normal; };

bool lp1(const PlaneF* pl, int planeNo, const V3& lpoint, const V3& ldir, float radius, const V3& opt, bool dirOpt, V3& result) {
  const float dp = dotf(lpoint, ldir);
  const float disc = sqr(dp) + sqr(radius) - absSq(lpoint);
  if (disc < 0.0f) return false;
  const float sq = std::sqrt(disc);
  float tLeft = -dp - sq, tRight = -dp + sq;
  for (int i = 0; i < planeNo; ++i) {
    const float numerator = dotf(sub(pl[i].point, lpoint), pl[i].normal);
    const float denominator = dotf(ldir, pl[i].normal);
    if (sqr(denominator) <= RVO_EPS) {
      if (numerator > 0.0f) return false;
      continue;
    }
    const float t = numerator / denominator;
    if (denominator >= 0.0f) tLeft = (tLeft > t) ? tLeft : t;   // windows.h max macro
    else tRight = (tRight < t) ? tRight : t;                    // windows.h min macro
    if (tLeft > tRight) return false;
  }
  if (dirOpt) {
    if (dotf(opt, ldir) > 0.0f) result = add(lpoint, mul(ldir, tRight));
    else result = add(lpoint, mul(ldir, tLeft));
  } else {
    const float t = dotf(ldir, sub(opt, lpoint));
    if (t < tLeft) result = add(lpoint, mul(ldir, tLeft));
    else if (t > tRight) result = add(lpoint, mul(ldir, tRight));
    else result = add(lpoint, mul(ldir, t));
  }
  return true;
}

bool lp2(const PlaneF* pl, int planeNo, float radius, const V3& opt, bool dirOpt, V3& result) {
  const float planeDist = dotf(pl[planeNo].point, pl[planeNo].normal);
  const float planeDistSq = sqr(planeDist);
  const float radiusSq = sqr(radius);
  if (planeDistSq > radiusSq) return false;
  const float planeRadiusSq = radiusSq - planeDistSq;
  const V3 planeCenter = mul(pl[planeNo].normal, planeDist);
  if (dirOpt) {
    const V3 pov = sub(opt, mul(pl[planeNo].normal, dotf(opt, pl[planeNo].normal)));
    const float lsq = absSq(pov);
    if (lsq <= RVO_EPS) result = planeCenter;
    else result = add(planeCenter, mul(pov, std::sqrt(planeRadiusSq / lsq)));
  } else {
    result = add(opt, mul(pl[planeNo].normal, dotf(sub(pl[planeNo].point, opt), pl[planeNo].normal)));
    if (absSq(result) > radiusSq) {
      const V3 pr = sub(result, planeCenter);
      const float prl = absSq(pr);
      result = add(planeCenter, mul(pr, std::sqrt(planeRadiusSq / prl)));
    }
  }
  for (int i = 0; i < planeNo; ++i) {
    if (dotf(pl[i].normal, sub(pl[i].point, result)) > 0.0f) {
      V3 cp = crossf(pl[i].normal, pl[planeNo].normal);
      if (absSq(cp) <= RVO_EPS) return false;
      const V3 ldir = normalize(cp);
      const V3 lnormal = crossf(ldir, pl[planeNo].normal);
      const V3 lpoint = add(pl[planeNo].point,
                            mul(lnormal, dotf(sub(pl[i].point, pl[planeNo].point), pl[i].normal) / dotf(lnormal, pl[i].normal)));
      if (!lp1(pl, i, lpoint, ldir, radius, opt, dirOpt, result)) return false;
    }
  }
  return true;
}

int lp3(const PlaneF* pl, int n, float radius, const V3& opt, bool dirOpt, V3& result) {
  if (dirOpt) result = mul(opt, radius);
  else if (absSq(opt) > sqr(radius)) result = mul(normalize(opt), radius);
  else result = opt;
  for (int i = 0; i < n; ++i) {
    if (dotf(pl[i].normal, sub(pl[i].point, result)) > 0.0f) {
      const V3 temp = result;
      if (!lp2(pl, i, radius, opt, dirOpt, result)) { result = temp; return i; }
    }
  }
  return n;
}

void lp4(const PlaneF* pl, int n, int beginPlane, float radius, V3& result) {
  float distance = 0.0f;
  std::vector<PlaneF> proj;
  for (int i = beginPlane; i < n; ++i) {
    if (dotf(pl[i].normal, sub(pl[i].point, result)) > distance) {
      proj.clear();
      for (int j = 0; j < i; ++j) {
        PlaneF p;
        const V3 cp = crossf(pl[j].normal, pl[i].normal);
        if (absSq(cp) <= RVO_EPS) {
          if (dotf(pl[i].normal, pl[j].normal) > 0.0f) continue;
          p.point = mul(add(pl[i].point, pl[j].point), 0.5f);
        } else {
          const V3 lnormal = crossf(cp, pl[i].normal);
          p.point = add(pl[i].point, mul(lnormal, dotf(sub(pl[j].point, pl[i].point), pl[j].normal) / dotf(lnormal, pl[j].normal)));
        }
        p.normal = normalize(sub(pl[j].normal, pl[i].normal));
        proj.push_back(p);
      }
      const V3 temp = result;
      if (lp3(proj.data(), (int)proj.size(), radius, pl[i].normal, true, result) < (int)proj.size()) result = temp;
      distance = dotf(pl[i].normal, sub(pl[i].point, result));
    }
  }
}

}  // namespace

// =================================================================================================
extern "C" void lqro_o_spheres(int P, double rx, double ry, double rz, double* pts) {
  double dlong = M_PI * (3.0 - std::sqrt(5.0));
  double dz = 2.0 / P;
  double longt = 0;
  double z = 1.0 - dz / 2.0;
  for (int i = 0; i < P; ++i) {
    pts[3 * i + 0] = 2 * rx * std::cos(longt) * std::sqrt(1 - z * z);
    pts[3 * i + 1] = 2 * ry * std::sin(longt) * std::sqrt(1 - z * z);
    pts[3 * i + 2] = 2 * rz * z;
    z -= dz;
    longt += dlong;
  }
}

extern "C" void lqro_o_horizon(const double* A, const double* B, const double* L, const double* E, int H, double* NF,
                               double* T, double* CG) {
  double BL[XD * XD], At[XD * XD], Bt[XD * 3], F[XD * XD], G[XD * 3], F2[XD * XD], G2[XD * 3];
  matmul(B, L, BL, XD, 4, XD);
  for (int i = 0; i < XD * XD; ++i) At[i] = A[i] + BL[i];  // Atilde = A + (B*L)
  matmul(B, E, Bt, XD, 4, 3);                               // Btilde = B*E
  for (int i = 0; i < XD * XD; ++i) F[i] = 0;
  for (int i = 0; i < XD; ++i) F[i * XD + i] = 1;
  for (int i = 0; i < XD * 3; ++i) G[i] = 0;
  for (int k = 0; k < H; ++k) {
    matmul(At, F, F2, XD, XD, XD);                          // F = Atilde*F
    matmul(At, G, G2, XD, XD, 3);                           // G = Atilde*G + Btilde
    for (int i = 0; i < XD * XD; ++i) F[i] = F2[i];
    for (int i = 0; i < XD * 3; ++i) G[i] = G2[i] + Bt[i];
    // (-P)*F with P = [I3 0]: row i of the product is -F(i,:) (the zero terms add +-0), MAIN:685
    for (int i = 0; i < 3; ++i) for (int j = 0; j < XD; ++j) NF[k * 48 + i * XD + j] = -F[i * XD + j];
    double cg[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cg[i * 3 + j] = G[i * 3 + j];  // P*G = rows 0..2 of G
    if (CG) for (int i = 0; i < 9; ++i) CG[k * 9 + i] = cg[i];
    inverse(cg, T + k * 9, 3);                              // Transform = !(P*G), MAIN:684
  }
}

extern "C" void lqro_o_candidates(const double* NF, const double* T, int H, int P, const double* sph, const double* xi,
                                  const double* xj, double* q) {
  double dx[XD];
  for (int i = 0; i < XD; ++i) dx[i] = xi[i] - xj[i];
  for (int k = 0; k < H; ++k) {
    double tr[3];
    for (int i = 0; i < 3; ++i) {
      double t = 0;
      for (int m = 0; m < XD; ++m) t += NF[k * 48 + i * XD + m] * dx[m];
      tr[i] = t;
    }
    const double* Tk = T + k * 9;
    for (int s = 0; s < P; ++s) {
      double v[3] = {sph[3 * s] + tr[0], sph[3 * s + 1] + tr[1], sph[3 * s + 2] + tr[2]};
      double* o = q + 3 * ((size_t)k * P + s);
      for (int i = 0; i < 3; ++i) {
        double t = 0;
        for (int m = 0; m < 3; ++m) t += Tk[i * 3 + m] * v[m];
        o[i] = t;
      }
    }
  }
}

extern "C" int lqro_o_reachable(const double* q, int n, const double* xi, const double* xj, double maxSpeed, int* idx) {
  double xr = maxSpeed, yr = maxSpeed, zr = maxSpeed;
  double xc = xi[3] - xj[3], yc = xi[4] - xj[4], zc = xi[5] - xj[5];
  int c = 0;
  for (int i = 0; i < n; ++i) {
    double a = q[3 * i] - xc, b = q[3 * i + 1] - yc, d = q[3 * i + 2] - zc;
    // pow(.,2) is an exact square (the compiler folds it to a multiply; verified against the
    // compiled reference in tests/test_oracle_vs_reference.py)
    if ((a * a) / (xr * xr) + (b * b) / (yr * yr) + (d * d) / (zr * zr) < 1.0) idx[c++] = i;
  }
  return c;
}

extern "C" double lqro_o_gjk(const double* vrel, const double* pts, int n, double* wpt_vrel, double* wpt_hull, int* stats) {
  double w1[3], w2[3];
  double r = gjk_point_cloud(vrel, pts, n, w1, w2, stats);
  for (int d = 0; d < 3; ++d) { if (wpt_vrel) wpt_vrel[d] = w1[d]; if (wpt_hull) wpt_hull[d] = w2[d]; }
  return r;
}

extern "C" int lqro_o_hull_escape(const double* pts, int n, const double* vrel, int fix_facet0_normal, double* distance,
                                  double* normal_io, int* best_facet, int* degenerate) {
  std::vector<double> r6(3 * (size_t)n);
  for (int i = 0; i < 3 * n; ++i) r6[i] = lqro_round6(pts[i]);      // MAIN:765-767 (6 significant digits)
  lqro_hull_result h;
  int rc = lqro_hull_qorder(r6.data(), n, &h);
  if (rc != 0) { lqro_hull_free(&h); return rc; }
  if (degenerate) *degenerate = h.degenerate;
  int best = 0;
  double best_d = 0;
  for (int f = 0; f < h.nfacets; ++f) {
    // the plane coefficients make the qhull 'n' text round trip ("%6.16g", MAIN:780-783)
    double nn[3];
    for (int k = 0; k < 3; ++k) {
      char buf[64];
      snprintf(buf, sizeof buf, "%6.16g", h.planes[4 * f + k]);
      nn[k] = strtod(buf, 0);
    }
    const double* p = pts + 3 * (size_t)h.verts[3 * f];             // first vertex, un-rounded cloud (MAIN:799-802)
    double d = std::fabs(nn[0] * (vrel[0] - p[0]) + nn[1] * (vrel[1] - p[1]) + nn[2] * (vrel[2] - p[2]));
    if (f == 0) {
      best_d = d;
      if (fix_facet0_normal) { normal_io[0] = nn[0]; normal_io[1] = nn[1]; normal_io[2] = nn[2]; }
    } else if (d < best_d) {
      best_d = d; best = f;
      normal_io[0] = nn[0]; normal_io[1] = nn[1]; normal_io[2] = nn[2];
    }
  }
  *distance = best_d;
  if (best_facet) *best_facet = best;
  int nf = h.nfacets;
  lqro_hull_free(&h);
  return nf;
}

extern "C" void lqro_o_pair(const double* NF, const double* T, int H, int P, const double* sph, const double* xi,
                            const double* xj, double maxSpeed, int fix_facet0_normal, const double* prev_normal,
                            lqro_o_pair_out* out) {
  memset(out, 0, sizeof *out);
  int n = H * P;
  std::vector<double> q(3 * (size_t)n), reach(3 * (size_t)n);
  std::vector<int> idx(n);
  lqro_o_candidates(NF, T, H, P, sph, xi, xj, q.data());
  int nr = lqro_o_reachable(q.data(), n, xi, xj, maxSpeed, idx.data());
  out->n_reach = nr;
  out->best_facet = -1;
  if (prev_normal) for (int d = 0; d < 3; ++d) out->normal[d] = prev_normal[d];
  if (nr <= 4) { out->status = 0; return; }                         // MAIN:1393
  for (int m = 0; m < nr; ++m) for (int d = 0; d < 3; ++d) reach[3 * m + d] = q[3 * (size_t)idx[m] + d];
  double vrel[3] = {xi[3] - xj[3], xi[4] - xj[4], xi[5] - xj[5]};
  double w1[3], w2[3];
  double sq = gjk_point_cloud(vrel, reach.data(), nr, w1, w2, out->gjk_stats);
  double distance = std::sqrt(sq);                                   // MAIN:741
  bool inside = distance < 0.0001 && distance > -0.0001;             // MAIN:752-757
  double normal[3] = {out->normal[0], out->normal[1], out->normal[2]};
  if (inside) {
    int nf = lqro_o_hull_escape(reach.data(), nr, vrel, fix_facet0_normal, &distance, normal, &out->best_facet, &out->degenerate);
    out->n_facets = nf;
    out->status = 2;
  } else {
    for (int d = 0; d < 3; ++d) normal[d] = (w1[d] - w2[d]) / distance;  // MAIN:743-745
    out->status = 1;
  }
  out->distance = distance;
  for (int d = 0; d < 3; ++d) out->normal[d] = normal[d];
  double half = distance * 0.5;                                      // MAIN:1401
  double mult = inside ? 1.0 : -1.0;                                 // MAIN:1088-1092
  for (int d = 0; d < 3; ++d) {
    out->plane[d] = (float)(xi[3 + d] + mult * half * normal[d]);   // MAIN:1094 (narrowed by Vector3(double,..), V3:113-118)
    out->plane[3 + d] = (float)normal[d];
  }
}

extern "C" void lqro_o_new_velocity(const float* planes, int n, const double* goalVel, double radius, double* newVel) {
  std::vector<PlaneF> pl(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) {
    pl[i].point = mk(planes[6 * i], planes[6 * i + 1], planes[6 * i + 2]);
    pl[i].normal = mk(planes[6 * i + 3], planes[6 * i + 4], planes[6 * i + 5]);
  }
  V3 pref = mk((float)goalVel[0], (float)goalVel[1], (float)goalVel[2]);
  V3 res = mk(0, 0, 0);
  float rad = (float)radius;
  int fail = lp3(pl.data(), n, rad, pref, false, res);
  if (fail < n) lp4(pl.data(), n, fail, rad, res);
  newVel[0] = res.x; newVel[1] = res.y; newVel[2] = res.z;
}

extern "C" void lqro_o_riccati(const double* A, const double* B, double qv, double qp, double r, int iters, double* L,
                               double* E, double* Lh, double* Eh) {
  // MAIN:492-495, 512-517.  V = [0 I3 0], Pmap = [I3 0]; Qx = 0, c = 0.
  const int n = XD;
  std::vector<double> V(3 * n, 0.0), Pm(3 * n, 0.0), Vt(n * 3), Pmt(n * 3), At(n * n), Bt(4 * n);
  V[0 * n + 3] = V[1 * n + 4] = V[2 * n + 5] = 1;
  Pm[0 * n + 0] = Pm[1 * n + 1] = Pm[2 * n + 2] = 1;
  transpose(V.data(), Vt.data(), 3, n);
  transpose(Pm.data(), Pmt.data(), 3, n);
  transpose(A, At.data(), n, n);
  transpose(B, Bt.data(), n, 4);
  std::vector<double> VtQV(n * n, 0.0), mVtQ(n * 3, 0.0);
  for (int i = 0; i < 3; ++i) { VtQV[(3 + i) * n + (3 + i)] = qv; mVtQ[(3 + i) * 3 + i] = -qv; }
  std::vector<double> S(VtQV), T(mVtQ), SB(n * 4), BtSB(16), M4(16), Minv(16), AtS(n * n), AtSB(n * 4), BtT(4 * 3), BtSA(4 * n),
      t1(n * 4), t2(n * 3), t3(n * n), AtT(n * 3), AtSA(n * n), SA(n * n);
  auto gain_core = [&](const std::vector<double>& Sx) {
    matmul(Sx.data(), B, SB.data(), n, n, 4);
    matmul(Bt.data(), SB.data(), BtSB.data(), 4, n, 4);
    for (int i = 0; i < 16; ++i) M4[i] = BtSB[i] + ((i % 5 == 0) ? r : 0.0);
    inverse(M4.data(), Minv.data(), 4);
  };
  for (int it = 0; it < iters; ++it) {
    gain_core(S);
    matmul(At.data(), S.data(), AtS.data(), n, n, n);
    matmul(AtS.data(), B, AtSB.data(), n, n, 4);
    matmul(AtSB.data(), Minv.data(), t1.data(), n, 4, 4);
    matmul(Bt.data(), T.data(), BtT.data(), 4, n, 3);
    matmul(t1.data(), BtT.data(), t2.data(), n, 4, 3);
    matmul(At.data(), T.data(), AtT.data(), n, n, 3);
    std::vector<double> Tn(n * 3);
    for (int i = 0; i < n * 3; ++i) Tn[i] = mVtQ[i] + AtT[i] - t2[i];
    matmul(S.data(), A, SA.data(), n, n, n);
    matmul(Bt.data(), SA.data(), BtSA.data(), 4, n, n);
    matmul(t1.data(), BtSA.data(), t3.data(), n, 4, n);
    matmul(AtS.data(), A, AtSA.data(), n, n, n);
    std::vector<double> Sn(n * n);
    for (int i = 0; i < n * n; ++i) Sn[i] = VtQV[i] + AtSA[i] - t3[i];
    T.swap(Tn); S.swap(Sn);
  }
  // MAIN:522-523
  gain_core(S);
  matmul(S.data(), A, SA.data(), n, n, n);
  matmul(Bt.data(), SA.data(), BtSA.data(), 4, n, n);
  matmul(Minv.data(), BtSA.data(), L, 4, 4, n);
  for (int i = 0; i < 4 * n; ++i) L[i] = -L[i];
  matmul(Bt.data(), T.data(), BtT.data(), 4, n, 3);
  matmul(Minv.data(), BtT.data(), E, 4, 4, 3);
  for (int i = 0; i < 12; ++i) E[i] = -E[i];
  // MAIN:526-530
  std::vector<double> Lt(n * 4), Et(3 * 4), RL(4 * n), RE(4 * 3), Qpt(n * n, 0.0), Rt(9), Pt(3 * n), Ati(n * n), Bti(n * 3), BL(n * n), tmp(n * n);
  transpose(L, Lt.data(), 4, n);
  transpose(E, Et.data(), 4, 3);
  for (int i = 0; i < 4 * n; ++i) RL[i] = r * L[i];
  for (int i = 0; i < 12; ++i) RE[i] = r * E[i];
  matmul(Lt.data(), RL.data(), tmp.data(), n, 4, n);
  for (int i = 0; i < n * n; ++i) Qpt[i] = 0.25 * tmp[i];
  for (int i = 0; i < 3; ++i) Qpt[i * n + i] += qp;
  matmul(Et.data(), RE.data(), Rt.data(), 3, 4, 3);
  for (int i = 0; i < 9; ++i) Rt[i] *= 0.25;
  matmul(Et.data(), RL.data(), Pt.data(), 3, 4, n);
  for (int i = 0; i < 3 * n; ++i) Pt[i] *= 0.25;
  matmul(B, L, BL.data(), n, 4, n);
  for (int i = 0; i < n * n; ++i) Ati[i] = A[i] + BL[i];
  matmul(B, E, Bti.data(), n, 4, 3);
  // MAIN:533-538
  std::vector<double> Atit(n * n), Btit(3 * n), Ptt(n * 3), mPQ(n * 3, 0.0);
  transpose(Ati.data(), Atit.data(), n, n);
  transpose(Bti.data(), Btit.data(), n, 3);
  transpose(Pt.data(), Ptt.data(), 3, n);
  for (int i = 0; i < 3; ++i) mPQ[i * 3 + i] = -qp;
  std::vector<double> St(Qpt), Tt(mPQ), SBt(n * 3), M3(9), M3i(9), AS(n * n), ASB(n * 3), Wm(n * 3), W1(n * 3), BT(9), W2(n * 3), AT(n * 3),
      SAt(n * n), BSA(3 * n), U(3 * n), W3(n * n), ASA(n * n);
  auto core3 = [&](const std::vector<double>& Sx) {
    matmul(Sx.data(), Bti.data(), SBt.data(), n, n, 3);
    matmul(Btit.data(), SBt.data(), M3.data(), 3, n, 3);
    for (int i = 0; i < 9; ++i) M3[i] += Rt[i];
    inverse(M3.data(), M3i.data(), 3);
  };
  for (int it = 0; it < iters; ++it) {
    core3(St);
    matmul(Atit.data(), St.data(), AS.data(), n, n, n);
    matmul(AS.data(), Bti.data(), ASB.data(), n, n, 3);
    for (int i = 0; i < n * 3; ++i) Wm[i] = Ptt[i] + ASB[i];
    matmul(Wm.data(), M3i.data(), W1.data(), n, 3, 3);
    matmul(Btit.data(), Tt.data(), BT.data(), 3, n, 3);
    matmul(W1.data(), BT.data(), W2.data(), n, 3, 3);
    matmul(Atit.data(), Tt.data(), AT.data(), n, n, 3);
    std::vector<double> Tn(n * 3);
    for (int i = 0; i < n * 3; ++i) Tn[i] = mPQ[i] + AT[i] - W2[i];
    matmul(St.data(), Ati.data(), SAt.data(), n, n, n);
    matmul(Btit.data(), SAt.data(), BSA.data(), 3, n, n);
    for (int i = 0; i < 3 * n; ++i) U[i] = Pt[i] + BSA[i];
    matmul(W1.data(), U.data(), W3.data(), n, 3, n);
    matmul(AS.data(), Ati.data(), ASA.data(), n, n, n);
    std::vector<double> Sn(n * n);
    for (int i = 0; i < n * n; ++i) Sn[i] = Qpt[i] + ASA[i] - W3[i];
    Tt.swap(Tn); St.swap(Sn);
  }
  // MAIN:542-543
  core3(St);
  matmul(St.data(), Ati.data(), SAt.data(), n, n, n);
  matmul(Btit.data(), SAt.data(), BSA.data(), 3, n, n);
  for (int i = 0; i < 3 * n; ++i) U[i] = Pt[i] + BSA[i];
  matmul(M3i.data(), U.data(), Lh, 3, 3, n);
  for (int i = 0; i < 3 * n; ++i) Lh[i] = -Lh[i];
  matmul(Btit.data(), Tt.data(), BT.data(), 3, n, 3);
  matmul(M3i.data(), BT.data(), Eh, 3, 3, 3);
  for (int i = 0; i < 9; ++i) Eh[i] = -Eh[i];
}

extern "C" void lqro_o_step(int N, int K, int H, int P, const double* NF, const double* T, const double* sph, const double* x,
                            const double* vpref, const int* nbr, double maxSpeed, double lpRadius, int fix_facet0_normal,
                            int nthreads, double* newV, float* planes_out, int* status_out, int* nplanes_out) {
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int t) {
    std::vector<float> planes(6 * (size_t)K);
    for (int i = t; i < N; i += nthreads) {
      int np = 0;
      for (int kk = 0; kk < K; ++kk) {
        int j = nbr[(size_t)i * K + kk];
        size_t slot = (size_t)i * K + kk;
        if (status_out) status_out[slot] = 0;
        if (planes_out) for (int d = 0; d < 6; ++d) planes_out[6 * slot + d] = 0.f;
        if (j < 0 || j == i) continue;
        lqro_o_pair_out po;
        lqro_o_pair(NF, T, H, P, sph, x + 16 * (size_t)i, x + 16 * (size_t)j, maxSpeed, fix_facet0_normal, 0, &po);
        if (status_out) status_out[slot] = po.status;
        if (po.status == 0) continue;
        for (int d = 0; d < 6; ++d) planes[6 * np + d] = po.plane[d];
        if (planes_out) for (int d = 0; d < 6; ++d) planes_out[6 * slot + d] = po.plane[d];
        ++np;
      }
      if (nplanes_out) nplanes_out[i] = np;
      lqro_o_new_velocity(planes.data(), np, vpref + 3 * (size_t)i, lpRadius, newV + 3 * (size_t)i);
    }
  };
  if (nthreads == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
  for (auto& t : th) t.join();
}

extern "C" void lqro_o_agent_update(int N, float* pos, float* vel, const float* newVel, float dt) {
  for (int i = 0; i < 3 * N; ++i) { vel[i] = newVel[i]; pos[i] += vel[i] * dt; }
}

// AGENT:153-174, restated: sorted insertion with a shrinking range
static void knn_rows(int N, const double* x, int begin, int end, int row0, int K, int* nbr) {
  std::vector<std::pair<double, int>> nb;
  for (int i = begin; i < end; ++i) {
    nb.clear();
    double rangeSq = std::numeric_limits<double>::infinity();
    const double* pi = x + (size_t)i * 16;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;                                     // AGENT:155
      const double* pj = x + (size_t)j * 16;
      const double d0 = pi[0] - pj[0], d1 = pi[1] - pj[1], d2 = pi[2] - pj[2];
      const double distSq = d0 * d0 + d1 * d1 + d2 * d2;        // absSq, V3:309-312
      if (distSq < rangeSq) {                                   // AGENT:158
        if ((int)nb.size() < K) nb.push_back(std::make_pair(distSq, j));
        size_t q = nb.size() - 1;
        while (q != 0 && distSq < nb[q - 1].first) { nb[q] = nb[q - 1]; --q; }
        nb[q] = std::make_pair(distSq, j);
        if ((int)nb.size() == K) rangeSq = nb.back().first;     // AGENT:169-171
      }
    }
    int* out = nbr + (size_t)(i - row0) * K;
    for (int q = 0; q < K; ++q) out[q] = q < (int)nb.size() ? nb[q].second : -1;
  }
}

extern "C" void lqro_o_knn(int N, const double* x, int begin, int n_local, int K, int nthreads, int* nbr) {
  if (K <= 0) return;
  if (nthreads <= 1) { knn_rows(N, x, begin, begin + n_local, begin, K, nbr); return; }
  std::vector<std::thread> th;
  const int per = (n_local + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    const int b = begin + t * per, e = std::min(begin + n_local, b + per);
    if (b < e) th.emplace_back(knn_rows, N, x, b, e, begin, K, nbr);
  }
  for (auto& t : th) t.join();
}
