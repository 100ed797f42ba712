// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// "Oracle by inclusion": this translation unit #includes the reference simulator source where it
// lies under /root/reference (nothing is copied into the repo) and exports a C ABI over the
// reference's own functions, so the tests can (1) pin oracle/lqro_oracle.cpp (our restatement)
// against the real code and (2) generate the golden vectors under tests/golden/.
// Built by oracle/Makefile into oracle/_ref/liblqro_ref.so; only exists where /root/reference does.
//
// What is reference code and what is harness code:
//   * every numerical routine called below (controlMatrices, findFG, createSpheres, createObstacle,
//     findReachableObstacle, pointInHull, run_gjk, convexHull, createHalfPlanes, calculateNewV,
//     riccatiControllerSteady(+Position), propagate, kalmanFilter1/2, sampleGaussian ...) is the
//     reference's, compiled unmodified;
//   * the loops in ref_pair()/ref_sim_avoid()/ref_sim_update() re-state the control flow of the
//     reference main loop (MAIN:1371-1418) with a run-time agent count, because the reference
//     hard-codes NUM_QUADS and keeps its state in locals of _tmain;
//   * the qconvex.exe spawn (MAIN:769-770, 814-815) is routed by oracle/shims/pre.h to
//     lqr_ref_system() below, which performs the same text-file round trip in-process using a
//     pluggable hull callback (default: oracle/hull_cpu.cpp; the tests can plug scipy/qhull).
#include "LQRObstaclesPositionControl(Working).cpp"

#include <cstring>
#include <cstdio>
#include <unistd.h>
#include <time.h>
#include "hull_cpu.h"

// ---------------------------------------------------------------- MSVC CRT rand() emulation
static unsigned int g_msvc_seed = 1u;
extern "C" int lqr_msvc_rand(void) {
  g_msvc_seed = g_msvc_seed * 214013u + 2531011u;
  return (int)((g_msvc_seed >> 16) & 0x7fffu);
}
extern "C" void lqr_msvc_srand(unsigned int seed) { g_msvc_seed = seed; }

// ---------------------------------------------------------------- qconvex stand-in
// callback: n points (as parsed from pointList.txt) -> triangular facets.
//   planes: 4 doubles per facet (unit normal, offset); verts: 3 point ids per facet (qhull Fv order).
// returns the number of facets, or <0 on failure.
typedef int (*lqr_hull_cb)(int n, const double* pts, int max_facets, double* planes, int* verts);
static lqr_hull_cb g_hull_cb = 0;
static std::vector<double> g_last_planes;
static std::vector<int> g_last_verts;
static int g_last_nfacets = -1;
static int g_last_npoints = 0;
static int g_hull_calls = 0, g_system_calls = 0;
static double g_hull_cb_s = 0;  // seconds inside the hull callback (the geometry alone, without the text round trip)
static inline double hull_now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static int default_hull_cb(int n, const double* pts, int max_facets, double* planes, int* verts) {
  lqro_hull_result r;
  int rc = lqro_hull_qorder(pts, n, &r);
  if (rc != 0) { lqro_hull_free(&r); return -1; }
  int nf = r.nfacets;
  if (nf > max_facets) { lqro_hull_free(&r); return -2; }
  for (int f = 0; f < nf; ++f) {
    for (int k = 0; k < 4; ++k) planes[4 * f + k] = r.planes[4 * f + k];
    for (int k = 0; k < 3; ++k) verts[3 * f + k] = r.verts[3 * f + k];
  }
  lqro_hull_free(&r);
  return nf;
}

extern "C" void ref_set_hull_callback(lqr_hull_cb cb) { g_hull_cb = cb; }
extern "C" int ref_last_hull(int* nfacets, int* npoints, double* planes, int* verts, int max_facets) {
  *nfacets = g_last_nfacets; *npoints = g_last_npoints;
  if (planes && verts && g_last_nfacets > 0) {
    int nf = g_last_nfacets < max_facets ? g_last_nfacets : max_facets;
    memcpy(planes, g_last_planes.data(), sizeof(double) * 4 * nf);
    memcpy(verts, g_last_verts.data(), sizeof(int) * 3 * nf);
  }
  return 0;
}
// seconds spent in the hull callback since the last reset: convexHull time (ref_stage_seconds()[3]) minus this is the cost of
// the text round trip of MAIN:762-806 (write pointList.txt, parse it, write and parse Planes.txt / facetVertices.txt), x4 per hull
extern "C" double ref_hull_callback_seconds(int reset) { const double v = g_hull_cb_s; if (reset) g_hull_cb_s = 0; return v; }
extern "C" void ref_counters(int* hull_calls, int* system_calls) { *hull_calls = g_hull_calls; *system_calls = g_system_calls; }

// Parses:  qconvex <n|Fv> TO "<out>" < "<in>"   and performs it with the hull callback.
// passthrough: hand the command line to the real shell (a `qconvex` on PATH does the work, e.g.
// tools/qconvex/qconvex = the GPU hull behind qhull's file formats) instead of the in-process stand-in
static int g_system_passthrough = 0;
extern "C" void ref_set_system_passthrough(int on) { g_system_passthrough = on; }
#undef system
extern "C" int lqr_ref_system(const char* cmd) {
  ++g_system_calls;
  if (g_system_passthrough) { ++g_hull_calls; return ::system(cmd); }
  char opt[16], outp[256], inp[256];
  if (sscanf(cmd, "qconvex %15s TO \"%255[^\"]\" < \"%255[^\"]\"", opt, outp, inp) != 3) return -1;
  FILE* fi = fopen(inp, "r");
  if (!fi) return 1;  // the reference's two stale-input calls on the very first hull (MAIN:814-815)
  int dim = 0, n = 0;
  if (fscanf(fi, "%d %d", &dim, &n) != 2 || dim != 3 || n < 4) { fclose(fi); return 1; }
  std::vector<double> pts(3 * (size_t)n);
  for (int i = 0; i < 3 * n; ++i) {
    if (fscanf(fi, "%lf", &pts[i]) != 1) { fclose(fi); return 1; }
  }
  fclose(fi);
  int maxf = 2 * n + 8;
  std::vector<double> planes(4 * (size_t)maxf);
  std::vector<int> verts(3 * (size_t)maxf);
  lqr_hull_cb cb = g_hull_cb ? g_hull_cb : default_hull_cb;
  const double tcb = hull_now_s();
  int nf = cb(n, pts.data(), maxf, planes.data(), verts.data());
  g_hull_cb_s += hull_now_s() - tcb;
  ++g_hull_calls;
  if (nf < 0) return 2;
  g_last_nfacets = nf; g_last_npoints = n;
  g_last_planes.assign(planes.begin(), planes.begin() + 4 * nf);
  g_last_verts.assign(verts.begin(), verts.begin() + 3 * nf);
  FILE* fo = fopen(outp, "w");
  if (!fo) return 3;
  if (strcmp(opt, "n") == 0) {
    // qhull 'n' format: dim+1, facet count, then "%6.16g " per coefficient (qh_REAL_1)
    fprintf(fo, "4\n%d\n", nf);
    for (int f = 0; f < nf; ++f) {
      for (int k = 0; k < 4; ++k) fprintf(fo, "%6.16g ", planes[4 * f + k]);
      fprintf(fo, "\n");
    }
  } else {
    // qhull 'Fv' format: facet count, then "<nverts> id id id"
    fprintf(fo, "%d\n", nf);
    for (int f = 0; f < nf; ++f) fprintf(fo, "3 %d %d %d\n", verts[3 * f], verts[3 * f + 1], verts[3 * f + 2]);
  }
  fclose(fo);
  return 0;
}

// ---------------------------------------------------------------- helpers
template <size_t R, size_t C> static void to_mat(const double* p, Matrix<R, C>& m) {
  for (size_t i = 0; i < R * C; ++i) m[i] = p[i];
}
template <size_t R, size_t C> static void from_mat(const Matrix<R, C>& m, double* p) {
  for (size_t i = 0; i < R * C; ++i) p[i] = m[i];
}

static bool g_inited = false;
extern "C" int ref_init(void) {
  if (g_inited) return 0;
  setup();  // MAIN:176-317 (constants + stubbed scene)
  // weights exactly as _tmain sets them, MAIN:1130-1141
  Qx = zeros<X_DIM, X_DIM>();
  Qv = zeros<V_DIM, V_DIM>();
  Qv(0, 0) = Qv(1, 1) = Qv(2, 2) = 100;
  Qp = zeros<V_DIM, V_DIM>();
  Qp(0, 0) = Qp(1, 1) = Qp(2, 2) = 0.1;
  R = 5 * identity<U_DIM>();
  M = 0.000000001 * identity<X_DIM>();
  N = 0.000000001 * identity<Z_DIM>();
  g_inited = true;
  return 0;
}
extern "C" void ref_set_weights(double qv, double qp, double r) {
  Qv = zeros<V_DIM, V_DIM>(); Qv(0, 0) = Qv(1, 1) = Qv(2, 2) = qv;
  Qp = zeros<V_DIM, V_DIM>(); Qp(0, 0) = Qp(1, 1) = Qp(2, 2) = qp;
  R = r * identity<U_DIM>();
}
extern "C" void ref_srand(unsigned int seed) { lqr_msvc_srand(seed); }
extern "C" int ref_num_points(void) { return NUM_POINTS; }
extern "C" int ref_obstacle_steps(void) { return OBSTACLE_STEPS; }
extern "C" double ref_nominal_input(void) { return nominalInput; }
extern "C" double ref_dt(void) { return dt; }
extern "C" int ref_chdir(const char* dir) { return chdir(dir); }

// MAIN:490-544
extern "C" void ref_control_matrices(const double* uGoal, const double* xstar, double* A, double* B, double* c,
                                     double* L, double* E, double* l, double* Lh, double* Eh) {
  Input u; to_mat(uGoal, u);
  Matrix<X_DIM, 1> xs; to_mat(xstar, xs);
  Matrix<X_DIM, X_DIM> mA; Matrix<X_DIM, U_DIM> mB; Matrix<X_DIM, 1> mc;
  Matrix<U_DIM, X_DIM> mL; Matrix<U_DIM, V_DIM> mE; Matrix<U_DIM, 1> ml;
  Matrix<V_DIM, X_DIM> mLh; Matrix<V_DIM, V_DIM> mEh;
  controlMatrices(u, xs, mA, mB, mc, mL, mE, ml, mLh, mEh);
  from_mat(mA, A); from_mat(mB, B); from_mat(mc, c); from_mat(mL, L); from_mat(mE, E); from_mat(ml, l);
  from_mat(mLh, Lh); from_mat(mEh, Eh);
}

// MAIN:425-440
extern "C" void ref_linearize_discretize(const double* x, const double* Rot, const double* u, double* A, double* B, double* c) {
  State mx; to_mat(x, mx); Rotation mR; to_mat(Rot, mR); Input mu; to_mat(u, mu);
  Matrix<X_DIM, X_DIM> mA; Matrix<X_DIM, U_DIM> mB; Matrix<X_DIM, 1> mc;
  linearizeDiscretize(mx, mR, mu, mA, mB, mc);
  from_mat(mA, A); from_mat(mB, B); from_mat(mc, c);
}

// MAIN:653-662
extern "C" void ref_find_fg(const double* A, const double* B, const double* L, const double* E, double* F, double* G) {
  Matrix<X_DIM, X_DIM> mA; Matrix<X_DIM, U_DIM> mB; Matrix<U_DIM, X_DIM> mL; Matrix<U_DIM, V_DIM> mE;
  Matrix<X_DIM, X_DIM> mF; Matrix<X_DIM, V_DIM> mG;
  to_mat(A, mA); to_mat(B, mB); to_mat(L, mL); to_mat(E, mE); to_mat(F, mF); to_mat(G, mG);
  findFG(mA, mB, mL, mE, mF, mG);
  from_mat(mF, F); from_mat(mG, G);
}

// MAIN:665-680
extern "C" int ref_create_spheres(double* pts) {
  std::vector<Matrix<3, 1> > points;
  createSpheres(points);
  for (size_t i = 0; i < points.size(); ++i) for (int k = 0; k < 3; ++k) pts[3 * i + k] = points[i][k];
  return (int)points.size();
}

// One ordered pair, control flow of MAIN:1381-1403.
//   normal_io : in = value normalVector holds before this pair (it persists across pairs in the
//               reference, MAIN:1365), out = value after.
//   returns status 0 (<=4 reachable, no plane), 1 (outside, GJK plane), 2 (inside, hull plane).
extern "C" int ref_pair(const double* A, const double* B, const double* L, const double* E, const double* xi,
                        const double* xj, int H, double maxSpeed, double* ellipsoids /*H*P*3 or 0*/,
                        int* reach_idx /*H*P or 0*/, int* n_reach, int* inside, double* distance_raw,
                        double* normal_io, float* plane /*6*/, int* gjk_stats /*6 or 0*/) {
  Matrix<X_DIM, X_DIM> mA; Matrix<X_DIM, U_DIM> mB; Matrix<U_DIM, X_DIM> mL; Matrix<U_DIM, V_DIM> mE;
  to_mat(A, mA); to_mat(B, mB); to_mat(L, mL); to_mat(E, mE);
  State x0, x1; to_mat(xi, x0); to_mat(xj, x1);
  Matrix<3, X_DIM> C = zeros<3, X_DIM>();
  C(0, 0) = C(1, 1) = C(2, 2) = 1;
  std::vector<Matrix<3, 1> > points; createSpheres(points);
  std::vector<Matrix<3, 1> > ell, reach;
  Matrix<3, 3> Transform; Matrix<3, 1> Translate;
  Matrix<X_DIM, X_DIM> Ft = identity<X_DIM>();
  Matrix<X_DIM, V_DIM> Gt = zeros<X_DIM, V_DIM>();
  size_t t = 0;
  for (int k = 0; k < H; ++k) {
    findFG(mA, mB, mL, mE, Ft, Gt);
    createObstacle(ell, points, Transform, Translate, t, Ft, Gt, C, x0, x1, 1);
  }
  findReachableObstacle(ell, reach, x0, x1, maxSpeed);
  if (ellipsoids) for (size_t i = 0; i < ell.size(); ++i) for (int k = 0; k < 3; ++k) ellipsoids[3 * i + k] = ell[i][k];
  if (reach_idx) {
    size_t m = 0;
    for (size_t i = 0; i < ell.size() && m < reach.size(); ++i)
      if (ell[i][0] == reach[m][0] && ell[i][1] == reach[m][1] && ell[i][2] == reach[m][2]) reach_idx[m++] = (int)i;
  }
  *n_reach = (int)reach.size();
  int status = 0;
  bool insideHull = false;
  double distance = 0;
  Matrix<3, 1> normalVector; to_mat(normal_io, normalVector);
  gjk_num_g_test = gjk_num_simplices = gjk_num_backups = gjk_num_dot_products = gjk_num_support_dp = gjk_num_other_ops = 0;
  if (reach.size() > 4) {
    pointInHull(reach, insideHull, x0, x1);
    if (gjk_stats) {
      gjk_stats[0] = gjk_num_g_test; gjk_stats[1] = gjk_num_simplices; gjk_stats[2] = gjk_num_backups;
      gjk_stats[3] = gjk_num_dot_products; gjk_stats[4] = gjk_num_support_dp; gjk_stats[5] = gjk_num_other_ops;
    }
    if (insideHull) { convexHull(reach, x0, x1, distance, normalVector); status = 2; }
    else { run_gjk(x0, x1, reach, distance, normalVector); status = 1; }
    *distance_raw = distance;
    distance *= 0.5;
    std::vector<Plane> planes;
    createHalfPlanes(x0, distance, normalVector, insideHull, planes);
    plane[0] = planes[0].point.x(); plane[1] = planes[0].point.y(); plane[2] = planes[0].point.z();
    plane[3] = planes[0].normal.x(); plane[4] = planes[0].normal.y(); plane[5] = planes[0].normal.z();
  } else if (gjk_stats) {
    for (int k = 0; k < 6; ++k) gjk_stats[k] = 0;
  }
  *inside = insideHull ? 1 : 0;
  from_mat(normalVector, normal_io);
  return status;
}

// MAIN:712-746 on an arbitrary cloud: point-vs-cloud GJK exactly as the reference calls it.
extern "C" void ref_run_gjk(const double* vrel, const double* pts, int n, double* distance, double* normal, int* gjk_stats) {
  State x0 = zeros<X_DIM>(), x1 = zeros<X_DIM>();
  x0[3] = vrel[0]; x0[4] = vrel[1]; x0[5] = vrel[2];
  std::vector<Matrix<3, 1> > reach(n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) reach[i][k] = pts[3 * i + k];
  Matrix<3, 1> nrm = zeros<3, 1>();
  gjk_num_g_test = gjk_num_simplices = gjk_num_backups = gjk_num_dot_products = gjk_num_support_dp = gjk_num_other_ops = 0;
  run_gjk(x0, x1, reach, *distance, nrm);
  from_mat(nrm, normal);
  if (gjk_stats) {
    gjk_stats[0] = gjk_num_g_test; gjk_stats[1] = gjk_num_simplices; gjk_stats[2] = gjk_num_backups;
    gjk_stats[3] = gjk_num_dot_products; gjk_stats[4] = gjk_num_support_dp; gjk_stats[5] = gjk_num_other_ops;
  }
}

// MAIN:809-838 on an arbitrary cloud (text round trip through the stand-in).
extern "C" void ref_convex_hull(const double* vrel, const double* pts, int n, double* distance, double* normal_io) {
  State x0 = zeros<X_DIM>(), x1 = zeros<X_DIM>();
  x0[3] = vrel[0]; x0[4] = vrel[1]; x0[5] = vrel[2];
  std::vector<Matrix<3, 1> > reach(n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) reach[i][k] = pts[3 * i + k];
  Matrix<3, 1> nrm; to_mat(normal_io, nrm);
  convexHull(reach, x0, x1, *distance, nrm);
  from_mat(nrm, normal_io);
}

// MAIN:1101-1112 (planes: n x 6 floats = point, normal)
extern "C" void ref_calculate_new_v(const float* planes, int n, const double* goalVel, double* newVel) {
  std::vector<Plane> pl(n);
  for (int i = 0; i < n; ++i) {
    pl[i].point = Vector3(planes[6 * i], planes[6 * i + 1], planes[6 * i + 2]);
    pl[i].normal = Vector3(planes[6 * i + 3], planes[6 * i + 4], planes[6 * i + 5]);
  }
  Matrix<3, 1> g, v; to_mat(goalVel, g);
  calculateNewV(pl, g, v);
  from_mat(v, newVel);
}

// ---------------------------------------------------------------- run-time-sized simulation (MAIN:1143-1418)
struct RefSim {
  int n;
  std::vector<Quadrotor*> q;
  Matrix<X_DIM, X_DIM> A; Matrix<X_DIM, U_DIM> B; Matrix<X_DIM, 1> c;
  std::vector<Matrix<3, 1> > points;
  // variables the reference declares outside its loops (MAIN:1360-1369) and so carries across pairs
  bool insideHull; double distance; Matrix<3, 1> normalVector;
  std::vector<Plane> orcaPlanes;
  size_t step;
};

extern "C" void* ref_sim_create(int n, const double* start_pos, const double* goal_pos, unsigned int seed) {
  ref_init();
  lqr_msvc_srand(seed);
  RefSim* s = new RefSim();
  s->n = n; s->step = 0;
  s->q.resize(n);
  Input uGoal; uGoal[0] = uGoal[1] = uGoal[2] = uGoal[3] = nominalInput;
  for (int i = 0; i < n; ++i) { s->q[i] = new Quadrotor(); s->q[i]->setupQuadVisualization(); }  // MAIN:1144-1147
  State xInit = zeros<X_DIM>();
  xInit[12] = xInit[13] = xInit[14] = xInit[15] = nominalInput;
  Rotation RotInit = identity<3>();
  Matrix<X_DIM, X_DIM> Pinit = 0.000000001 * identity<X_DIM>();
  State xGoal = zeros<X_DIM>();
  xGoal[12] = xGoal[13] = xGoal[14] = xGoal[15] = nominalInput;
  Rotation RotGoal = identity<3>();
  for (int i = 0; i < n; ++i) {  // pattern of MAIN:1175-1179 per quad
    xInit[0] = start_pos[3 * i]; xInit[1] = start_pos[3 * i + 1]; xInit[2] = start_pos[3 * i + 2];
    xInit[3] = 0.0; xInit[4] = 0.0; xInit[5] = 0.0;
    Position pg; pg[0] = goal_pos[3 * i]; pg[1] = goal_pos[3 * i + 1]; pg[2] = goal_pos[3 * i + 2];
    s->q[i]->setupQuadrotors(xInit, RotInit, Pinit, xGoal, pg, uGoal, RotGoal);
  }
  for (int i = 0; i < n; ++i) {  // MAIN:1353-1356
    s->q[i]->findMatrices(s->A, s->B, s->c);
    s->q[i]->vGoal = s->q[i]->findVGoal();
  }
  createSpheres(s->points);  // MAIN:1359
  s->insideHull = false; s->distance = 0; s->normalVector = zeros<3, 1>();
  return s;
}
extern "C" void ref_sim_destroy(void* h) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) delete s->q[i];
  delete s;
}
extern "C" void ref_sim_get(void* h, double* x, double* xTrue, double* vGoal, double* Rot, double* RotTrue, double* P) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) {
    if (x) from_mat(s->q[i]->x, x + 16 * i);
    if (xTrue) from_mat(s->q[i]->xTrue, xTrue + 16 * i);
    if (vGoal) from_mat(s->q[i]->vGoal, vGoal + 3 * i);
    if (Rot) from_mat(s->q[i]->Rot, Rot + 9 * i);
    if (RotTrue) from_mat(s->q[i]->RotTrue, RotTrue + 9 * i);
    if (P) from_mat(s->q[i]->P, P + 256 * i);
  }
}
extern "C" void ref_sim_get_dynamics(void* h, double* A, double* B, double* L, double* E, double* Lh, double* Eh) {
  RefSim* s = (RefSim*)h;
  from_mat(s->A, A); from_mat(s->B, B);
  from_mat(s->q[0]->L, L); from_mat(s->q[0]->E, E); from_mat(s->q[0]->Lh, Lh); from_mat(s->q[0]->Eh, Eh);
}
extern "C" void ref_sim_set_x(void* h, const double* x) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) to_mat(x + 16 * i, s->q[i]->x);
}
extern "C" void ref_sim_set_vgoal(void* h, const double* v) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) to_mat(v + 3 * i, s->q[i]->vGoal);
}

// Avoidance half of a time step: MAIN:1373-1407 with run-time n and H.
//   planes_out : n*(n-1)*6 floats or 0;  status_out : n*(n-1) ints or 0 (0 none, 1 outside, 2 inside)
//   nplanes_out: n ints or 0.  Slot (i, jj) with jj = j - (j>i).
static int g_fix_facet0 = 0;       // harness option, see ref_sim_set_fix_facet0
static long g_facet0_events = 0;
// The reference leaves `normal` untouched when hull facet 0 is the nearest one (MAIN:825-836), so the plane
// inherits the previous pair's normal.  With this option the HARNESS (not the reference code) overwrites the
// normal with facet 0's after convexHull returns, which is what a pair-parallel implementation computes.
// Either way the number of such events is counted.
extern "C" void ref_sim_set_fix_facet0(int on) { g_fix_facet0 = on; }
extern "C" long ref_sim_facet0_events(void) { return g_facet0_events; }

extern "C" void ref_sim_avoid(void* h, int H, double maxSpeed, double* newV, float* planes_out, int* status_out,
                              int* nplanes_out, int* nreach_out) {
  RefSim* s = (RefSim*)h;
  int n = s->n;
  Matrix<3, X_DIM> C = zeros<3, X_DIM>();
  C(0, 0) = C(1, 1) = C(2, 2) = 1;
  Matrix<X_DIM, X_DIM> Ft; Matrix<X_DIM, V_DIM> Gt;
  size_t t = s->step;
  for (int i = 0; i < n; ++i) {
    Quadrotor* qi = s->q[i];
    for (int j = 0; j < n; ++j) {
      qi->ellipsoids.clear();
      qi->reachablePoints.clear();
      if (i == j) continue;
      int slot = i * (n - 1) + (j - (j > i ? 1 : 0));
      Ft = identity<X_DIM>();
      Gt = zeros<X_DIM, V_DIM>();
      for (int k = 0; k < H; ++k) {
        findFG(s->A, s->B, qi->L, qi->E, Ft, Gt);
        createObstacle(qi->ellipsoids, s->points, qi->Transform, qi->Translate, t, Ft, Gt, C, qi->x, s->q[j]->x, 1);
      }
      findReachableObstacle(qi->ellipsoids, qi->reachablePoints, qi->x, s->q[j]->x, maxSpeed);
      int numObstacle = (int)qi->reachablePoints.size();
      if (nreach_out) nreach_out[slot] = numObstacle;
      int st = 0;
      if (numObstacle > 4) {
        pointInHull(qi->reachablePoints, s->insideHull, qi->x, s->q[j]->x);
        if (s->insideHull) {
          convexHull(qi->reachablePoints, qi->x, s->q[j]->x, s->distance, s->normalVector); st = 2;
          if (g_last_nfacets > 0) {  // was facet 0 the nearest?  (distance is then exactly facet 0's)
            const int v0 = g_last_verts[0];
            const double d0 = abs(g_last_planes[0] * ((qi->x[3] - s->q[j]->x[3]) - qi->reachablePoints[v0][0]) +
                                  g_last_planes[1] * ((qi->x[4] - s->q[j]->x[4]) - qi->reachablePoints[v0][1]) +
                                  g_last_planes[2] * ((qi->x[5] - s->q[j]->x[5]) - qi->reachablePoints[v0][2]));
            bool nearest0 = true;
            for (int f = 1; f < g_last_nfacets && nearest0; ++f) {
              const int vf = g_last_verts[3 * f];
              const double df = abs(g_last_planes[4 * f] * ((qi->x[3] - s->q[j]->x[3]) - qi->reachablePoints[vf][0]) +
                                    g_last_planes[4 * f + 1] * ((qi->x[4] - s->q[j]->x[4]) - qi->reachablePoints[vf][1]) +
                                    g_last_planes[4 * f + 2] * ((qi->x[5] - s->q[j]->x[5]) - qi->reachablePoints[vf][2]));
              if (df < d0) nearest0 = false;
            }
            if (nearest0) {
              ++g_facet0_events;
              if (g_fix_facet0) { s->normalVector[0] = g_last_planes[0]; s->normalVector[1] = g_last_planes[1]; s->normalVector[2] = g_last_planes[2]; }
            }
          }
        }
        else { run_gjk(qi->x, s->q[j]->x, qi->reachablePoints, s->distance, s->normalVector); st = 1; }
        s->distance *= 0.5;
        createHalfPlanes(qi->x, s->distance, s->normalVector, s->insideHull, s->orcaPlanes);
        if (planes_out) {
          const Plane& p = s->orcaPlanes.back();
          float* o = planes_out + 6 * (size_t)slot;
          o[0] = p.point.x(); o[1] = p.point.y(); o[2] = p.point.z();
          o[3] = p.normal.x(); o[4] = p.normal.y(); o[5] = p.normal.z();
        }
      }
      if (status_out) status_out[slot] = st;
    }
    if (nplanes_out) nplanes_out[i] = (int)s->orcaPlanes.size();
    calculateNewV(s->orcaPlanes, qi->vGoal, qi->newV);
    if (newV) from_mat(qi->newV, newV + 3 * i);
  }
}
extern "C" void ref_sim_set_newv(void* h, const double* newV) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) to_mat(newV + 3 * i, s->q[i]->newV);
}
// Update half of a time step: MAIN:1408-1417.
extern "C" void ref_sim_update(void* handle) {
  RefSim* s = (RefSim*)handle;
  for (int i = 0; i < s->n; ++i) {
    Quadrotor* q = s->q[i];
    q->vGoal = q->newV;
    Input u = q->findU();
    q->propagateU(u);
    kalmanFilter1(q->x, q->Rot, u, M, q->P);
    Observation z = sampleGaussian(::h(q->xTrue, q->RotTrue), N);
    kalmanFilter2(q->x, q->Rot, z, N, q->P);
    q->visualize(s->step * dt);
    q->vGoal = q->findVGoal();
  }
  s->step++;
}

// Avoidance step of a SUBSET of agents with explicit neighbour lists (the synthetic swarms): the loop body
// of MAIN:1376-1406 per listed agent, all in the reference's own functions.  x: n_all*16 estimates,
// vpref: n_all*3, nbr_rows: n_sel*K neighbour ids of the selected agents (<0 = empty), newV: n_sel*3.
// Returns the number of (i,j) pairs evaluated.  This is what the CPU baseline times.
// wall-clock seconds spent in the reference's own three categories ("obstacle creation" = findFG + createObstacle +
// findReachableObstacle, "half-plane creation" = pointInHull + run_gjk/convexHull + createHalfPlanes, "new velocity" =
// calculateNewV; the category names of the reference's timing printouts) plus the hull branch alone, accumulated by
// ref_step_agents since the last ref_stage_seconds(reset = 1)
static double g_stage_s[4] = {0, 0, 0, 0};
static long g_stage_hulls = 0;
static inline double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
extern "C" void ref_stage_seconds(double* out4, long* hulls, int reset) {
  if (out4) for (int q = 0; q < 4; ++q) out4[q] = g_stage_s[q];
  if (hulls) *hulls = g_stage_hulls;
  if (reset) { for (int q = 0; q < 4; ++q) g_stage_s[q] = 0; g_stage_hulls = 0; }
}

extern "C" long ref_step_agents(const double* A, const double* B, const double* L, const double* E, const double* x,
                                const double* vpref, const int* nbr_rows, const int* agents, int n_sel, int K, int H,
                                double maxSpeed, double* newV, int* nplanes_out) {
  Matrix<X_DIM, X_DIM> mA; Matrix<X_DIM, U_DIM> mB; Matrix<U_DIM, X_DIM> mL; Matrix<U_DIM, V_DIM> mE;
  to_mat(A, mA); to_mat(B, mB); to_mat(L, mL); to_mat(E, mE);
  Matrix<3, X_DIM> C = zeros<3, X_DIM>();
  C(0, 0) = C(1, 1) = C(2, 2) = 1;
  std::vector<Matrix<3, 1> > points; createSpheres(points);
  std::vector<Matrix<3, 1> > ell, reach;
  ell.reserve(NUM_POINTS * OBSTACLE_STEPS); reach.reserve(NUM_POINTS * OBSTACLE_STEPS);
  Matrix<3, 3> Transform; Matrix<3, 1> Translate;
  Matrix<X_DIM, X_DIM> Ft; Matrix<X_DIM, V_DIM> Gt;
  bool insideHull = false; double distance = 0; Matrix<3, 1> normalVector = zeros<3, 1>();
  std::vector<Plane> orcaPlanes;
  size_t t = 0;
  long npairs = 0;
  for (int a = 0; a < n_sel; ++a) {
    const int i = agents[a];
    State x0; to_mat(x + 16 * (size_t)i, x0);
    for (int kk = 0; kk < K; ++kk) {
      const int j = nbr_rows[(size_t)a * K + kk];
      ell.clear(); reach.clear();
      if (j < 0 || j == i) continue;
      ++npairs;
      State x1; to_mat(x + 16 * (size_t)j, x1);
      const double t0 = now_s();
      Ft = identity<X_DIM>(); Gt = zeros<X_DIM, V_DIM>();
      for (int k = 0; k < H; ++k) {
        findFG(mA, mB, mL, mE, Ft, Gt);
        createObstacle(ell, points, Transform, Translate, t, Ft, Gt, C, x0, x1, 1);
      }
      findReachableObstacle(ell, reach, x0, x1, maxSpeed);
      const double t1 = now_s();
      g_stage_s[0] += t1 - t0;
      if ((int)reach.size() > 4) {
        pointInHull(reach, insideHull, x0, x1);
        if (insideHull) {
          const double th = now_s();
          convexHull(reach, x0, x1, distance, normalVector);
          g_stage_s[3] += now_s() - th;
          ++g_stage_hulls;
        } else run_gjk(x0, x1, reach, distance, normalVector);
        distance *= 0.5;
        createHalfPlanes(x0, distance, normalVector, insideHull, orcaPlanes);
        g_stage_s[1] += now_s() - t1;
      }
    }
    if (nplanes_out) nplanes_out[a] = (int)orcaPlanes.size();
    Matrix<3, 1> g, v; to_mat(vpref + 3 * (size_t)i, g);
    const double t2 = now_s();
    calculateNewV(orcaPlanes, g, v);
    g_stage_s[2] += now_s() - t2;
    from_mat(v, newV + 3 * (size_t)a);
  }
  return npairs;
}

// The unmodified reference main() (renamed by shims/tchar.h).  stdin must not block: the reference
// ends with `std::cin >> k` (MAIN:1421-1422); callers redirect stdin from /dev/null.
extern "C" int ref_run_tmain(void) {
  g_inited = false;  // _tmain calls setup() itself and sets the weights
  lqr_msvc_srand(1u);
  int rc = ref_tmain(0, 0);
  g_inited = true;   // setup() ran and the weights are the _tmain ones
  return rc;
}

// ---------------------------------------------------------------- simulator-form update parity
extern "C" unsigned int ref_get_seed(void) { return g_msvc_seed; }
// `count` draws of the reference's own normal() (MAIN:325-336) from the current stream position
extern "C" void ref_normals(int count, double* out) { for (int i = 0; i < count; ++i) out[i] = normal(); }
// full per-agent state in / out (any pointer may be 0)
extern "C" void ref_sim_set_state(void* h, const double* x, const double* xTrue, const double* Rot, const double* RotTrue,
                                  const double* P) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) {
    if (x) to_mat(x + 16 * i, s->q[i]->x);
    if (xTrue) to_mat(xTrue + 16 * i, s->q[i]->xTrue);
    if (Rot) to_mat(Rot + 9 * i, s->q[i]->Rot);
    if (RotTrue) to_mat(RotTrue + 9 * i, s->q[i]->RotTrue);
    if (P) to_mat(P + 256 * i, s->q[i]->P);
  }
}
extern "C" void ref_sim_get_goals(void* h, double* pGoal, double* newV) {
  RefSim* s = (RefSim*)h;
  for (int i = 0; i < s->n; ++i) {
    if (pGoal) from_mat(s->q[i]->pGoal, pGoal + 3 * i);
    if (newV) from_mat(s->q[i]->newV, newV + 3 * i);
  }
}
// The reference's globals and quad 0's gains in the layout of lqro_sim_model (include/lqro.h): 458 doubles
extern "C" int ref_sim_model(void* h, double* out) {
  RefSim* s = (RefSim*)h;
  double* p = out;
  *p++ = dt; *p++ = gravity; *p++ = mass; *p++ = momentConst; *p++ = thrust_latency; *p++ = length_; *p++ = jStep;
  from_mat(inertia, p); p += 9;
  from_mat(invInertia, p); p += 9;
  from_mat(s->q[0]->uGoal, p); p += 4;
  from_mat(s->q[0]->L, p); p += 64;
  from_mat(s->q[0]->E, p); p += 12;
  from_mat(s->q[0]->l, p); p += 4;
  from_mat(s->q[0]->Lh, p); p += 48;
  from_mat(s->q[0]->Eh, p); p += 9;
  from_mat(M, p); p += 256;
  from_mat(N, p); p += 36;
  return (int)(p - out);
}
// Stage-level entry points on explicit arguments (unit parity of the pieces of the update)
extern "C" void ref_exp16(const double* q, double* out) { Matrix<16, 16> m; to_mat(q, m); from_mat(exp(m), out); }
extern "C" void ref_exp3(const double* q, double* out) { Matrix<3, 3> m; to_mat(q, m); from_mat(exp(m), out); }
extern "C" void ref_jacobi16(const double* q, double* V, double* D) {
  Matrix<16, 16> m, v, d; to_mat(q, m); jacobi(m, v, d); from_mat(v, V); from_mat(d, D);
}
extern "C" void ref_f(const double* x, const double* Rot, const double* u, double* xdot) {
  State xs; Rotation R; Input us; to_mat(x, xs); to_mat(Rot, R); to_mat(u, us); from_mat(f(xs, R, us), xdot);
}
