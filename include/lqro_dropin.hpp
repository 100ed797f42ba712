// include/lqro_dropin.hpp -- C++ shims carrying the REFERENCE'S OWN SIGNATURES for the hot path, so
// that the reference main loop (MAIN:1371-1418, or a headless copy of it) links against the B200 path
// instead of its CPU functions.  Header-only; calls nothing but the C ABI of include/lqro.h.
//
// The shims are templates over the reference's `Matrix<R,C>` (include/matrix.h: value type, row-major
// `double _elems[R*C]`, operator[] / operator()), so this header does not need -- and does not copy --
// the reference's matrix header: include it first and instantiate with its types.  `PlaneT` is the
// reference's `struct Plane { Vector3 point, normal; }` (MAIN:844-855), `Vec3T` its Vector3 (FP32).
//
// Conventions preserved from the reference: everything returns void, outputs by non-const
// reference, std::vector outputs are APPENDED to (callers clear them, MAIN:1377-1378), calculateNewV
// clears its input vector (MAIN:1111), no exceptions; a failing device call terminates the process
// the way the reference's assert/exit(-1) paths do.
//
// Two levels:
//   * per-function shims below: same names, same argument lists (createObstacle ... calculateNewV,
//     controlMatrices) -- each call is one small device round trip: correct, slow;
//   * lqro::Swarm -- the whole `for i / for j` block of MAIN:1373-1407 for all agents in ONE device
//     step (lqro_step): this is what replaces the hot loop; Swarm::update replaces the second loop of the
//     time step (MAIN:1408-1417: LQR feedback, propagate, Kalman filter) with one lqro_sim_update call.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lqro.h"

namespace lqro {

// The reference's Matrix::operator[] const returns by value (MAT:58-61), so matrices are copied out element-wise.
template <class M>
inline std::vector<double> flat(const M& m, int n) {
  std::vector<double> v(n);
  for (int i = 0; i < n; ++i) v[i] = m[i];
  return v;
}

inline void check(int rc, lqro_ctx* ctx, const char* what) {
  if (rc != LQRO_OK) {
    std::fprintf(stderr, "liblqro: %s failed (%d): %s\n", what, rc, lqro_last_error(ctx));
    std::exit(-1);  // the reference's own failure mode (MAT:859,870,974,1104)
  }
}

// One process-wide context for the per-function shims, created on first use from the reference's
// compile-time constants.  H and P follow OBSTACLE_STEPS / NUM_POINTS if those macros are visible.
inline lqro_ctx*& shim_ctx() {
  static lqro_ctx* ctx = nullptr;
  return ctx;
}
inline int& shim_hp() { static int hp = 0; return hp; }        // H*P of the shim context
inline int& shim_samples() { static int p = 0; return p; }   // P of the shim context
inline lqro_ctx* ensure_ctx() {
  lqro_ctx*& ctx = shim_ctx();
  if (!ctx) {
    lqro_config cfg;
    lqro_default_config(&cfg);
#ifdef OBSTACLE_STEPS
    cfg.horizon = OBSTACLE_STEPS;
#endif
#ifdef NUM_POINTS
    cfg.n_samples = NUM_POINTS;
#endif
#ifdef XYRADIUS
    cfg.xy_radius = XYRADIUS;
    cfg.z_radius = ZRADIUS;
#endif
    cfg.n_agents = 2;
    cfg.max_neighbors = 1;
    check(lqro_create(&cfg, &ctx), nullptr, "lqro_create");
    shim_hp() = cfg.horizon * cfg.n_samples;
    shim_samples() = cfg.n_samples;
  }
  return ctx;
}

// ---- MAIN:653-662 findFG(A, B, L, E, F, G): the reference's own argument list and meaning, one device call per call.
// F <- (A + B L) F, G <- (A + B L) G + B E, in place, in the reference's operation order (same bits).
template <class M16x16, class M16x4, class M4x16, class M4x3, class M16x3>
inline void findFG(const M16x16& A, const M16x4& B, const M4x16& L, const M4x3& E, M16x16& F, M16x3& G) {
  lqro_ctx* ctx = ensure_ctx();
  std::vector<double> f = flat(F, 256), g = flat(G, 48);
  check(lqro_find_fg(ctx, flat(A, 256).data(), flat(B, 64).data(), flat(L, 64).data(), flat(E, 12).data(), f.data(), g.data()), ctx, "lqro_find_fg");
  for (int i = 0; i < 256; ++i) F[i] = f[i];
  for (int i = 0; i < 48; ++i) G[i] = g[i];
}

// The H findFG calls the reference makes PER PAIR (MAIN:1383-1384) compute pair-independent tables: setDynamics builds all of
// them once on the device (lqro_set_dynamics); lqro::Swarm::step and createObstacleAll use them.
template <class M16x16, class M16x4, class M4x16, class M4x3>
inline void setDynamics(const M16x16& A, const M16x4& B, const M4x16& L, const M4x3& E) {
  lqro_ctx* ctx = ensure_ctx();
  check(lqro_set_dynamics(ctx, flat(A, 256).data(), flat(B, 64).data(), flat(L, 64).data(), flat(E, 12).data()), ctx, "lqro_set_dynamics");
}

// ---- MAIN:490-544 controlMatrices, Riccati part (A, B already linearised/discretised by the caller:
// the quadrotor model MAIN:338-440 stays host code).  Fills L, E, Lh, Eh; l = 0 for c = 0, Qx = 0.
template <class M16x16, class M16x4, class M4x16, class M4x3, class M3x16, class M3x3>
inline void controlMatricesRiccati(const M16x16& A, const M16x4& B, double qv, double qp, double r, M4x16& L, M4x3& E, M3x16& Lh,
                                   M3x3& Eh) {
  lqro_ctx* ctx = ensure_ctx();
  double l[64], e[12], lh[48], eh[9];
  check(lqro_solve_gains(ctx, 1, flat(A, 256).data(), flat(B, 64).data(), qv, qp, r, 300, l, e, lh, eh), ctx, "lqro_solve_gains");
  for (int i = 0; i < 64; ++i) L[i] = l[i];
  for (int i = 0; i < 12; ++i) E[i] = e[i];
  for (int i = 0; i < 48; ++i) Lh[i] = lh[i];
  for (int i = 0; i < 9; ++i) Eh[i] = eh[i];
}

// ---- MAIN:490-544 controlMatrices with the reference's own argument list (SIM2:43): linearizeDiscretize about the
// hover point (xHat = [0 .. 0, uGoal], RHat = I, MAIN:497-509) and both Riccati recursions on the device.
// `xstar` is unused, as in the reference.  Weights Qv = qv*I, Qp = qp*I, R = r*I are the reference's globals
// (MAIN:1130-1137: 100, 0.1, 5); the physical constants are setup()'s (lqro_sim_default_model) unless `model` is given.
template <class Input, class M16, class M16x16, class M16x4, class M4x16, class M4x3, class M4, class M3x16, class M3x3>
inline void controlMatrices(const Input& uGoal, const M16& /*xstar*/, M16x16& A, M16x4& B, M16& c, M4x16& L, M4x3& E, M4& l, M3x16& Lh,
                            M3x3& Eh, double qv = 100.0, double qp = 0.1, double r = 5.0, const lqro_sim_model* model = nullptr) {
  lqro_ctx* ctx = ensure_ctx();
  lqro_sim_model m;
  if (model) m = *model; else lqro_sim_default_model(&m);
  check(lqro_sim_set_model(ctx, &m), ctx, "lqro_sim_set_model");
  double x[16] = {0}, Rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, u[4], a[256], b[64], cc[16], ll[64], e[12], lh[48], eh[9];
  for (int i = 0; i < 4; ++i) { u[i] = uGoal[i]; x[12 + i] = uGoal[i]; }
  check(lqro_linearize_discretize(ctx, 1, x, Rot, u, a, b, cc), ctx, "lqro_linearize_discretize");
  check(lqro_solve_gains(ctx, 1, a, b, qv, qp, r, 300, ll, e, lh, eh), ctx, "lqro_solve_gains");
  for (int i = 0; i < 256; ++i) A[i] = a[i];
  for (int i = 0; i < 64; ++i) { B[i] = b[i]; L[i] = ll[i]; }
  for (int i = 0; i < 16; ++i) c[i] = cc[i];
  for (int i = 0; i < 12; ++i) E[i] = e[i];
  // l = -(R + B'SB)^-1 (B'S c + B'a), a = pinv(..)(Qx xstar - A'S c + ..) (MAIN:519-524): with Qx = 0 and c = 0 -- the hover
  // point is an equilibrium, f = 0 -- both vanish, l = 0 (the reference's value).  Any other point is refused.
  for (int i = 0; i < 16; ++i)
    if (cc[i] != 0.0) { std::fprintf(stderr, "lqro::controlMatrices: c != 0 (not an equilibrium): l is not available\n"); std::exit(-1); }
  for (int i = 0; i < 4; ++i) l[i] = 0.0;
  for (int i = 0; i < 48; ++i) Lh[i] = lh[i];
  for (int i = 0; i < 9; ++i) Eh[i] = eh[i];
}

// ---- MAIN:665-680 createSpheres(std::vector<Matrix<3,1>>& points): appends NUM_POINTS samples of the doubled ellipsoid
template <class V3d>
inline void createSpheres(std::vector<V3d>& points) {
  lqro_ctx* ctx = ensure_ctx();
  const int P = shim_samples();
  std::vector<double> sph(3 * (size_t)P);
  check(lqro_get_spheres(ctx, sph.data()), ctx, "lqro_get_spheres");
  for (int i = 0; i < P; ++i) {
    V3d p;
    p[0] = sph[3 * i]; p[1] = sph[3 * i + 1]; p[2] = sph[3 * i + 2];
    points.push_back(p);
  }
}

// ---- MAIN:683-688 createObstacle(ellipsoids, points, Transform, Translate, t, F, G, P, xInit1, xInit2, radius): the reference's
// own argument list (`t` and `radius` are unused there too).  Transform = !(P*G), Translate = -P*F*(xInit1-xInit2), appends
// Transform*(points[i] + Translate) for every sample.
template <class V3d, class M3x3, class M16x16, class M16x3, class M3x16, class State>
inline void createObstacle(std::vector<V3d>& ellipsoids, const std::vector<V3d>& points, M3x3& Transform, V3d& Translate, const size_t& /*t*/,
                           const M16x16& F, const M16x3& G, const M3x16& P, const State& xInit1, const State& xInit2, const int& /*radius*/) {
  lqro_ctx* ctx = ensure_ctx();
  const int n = (int)points.size();
  std::vector<double> p(3 * (size_t)(n > 0 ? n : 1)), q(3 * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) p[3 * i + k] = points[i][k];
  double T[9], tr[3];
  check(lqro_create_obstacle(ctx, n, p.data(), flat(F, 256).data(), flat(G, 48).data(), flat(P, 48).data(), flat(xInit1, 16).data(),
                             flat(xInit2, 16).data(), T, tr, q.data()), ctx, "lqro_create_obstacle");
  for (int i = 0; i < 9; ++i) Transform[i] = T[i];
  for (int i = 0; i < 3; ++i) Translate[i] = tr[i];
  for (int i = 0; i < n; ++i) {
    V3d v;
    v[0] = q[3 * i]; v[1] = q[3 * i + 1]; v[2] = q[3 * i + 2];
    ellipsoids.push_back(v);
  }
}

// createObstacle for ALL horizon steps of one pair from the tables of setDynamics: appends H*P points (k-major), i.e. what the
// reference's `for k` loop around findFG + createObstacle (MAIN:1383-1389) leaves in `ellipsoids`, in one device call.
// H and P are the context's (the configuration the shim context was created with), not re-read from macros here.
template <class V3d, class State>
inline void createObstacleAll(std::vector<V3d>& ellipsoids, const State& xInit1, const State& xInit2) {
  lqro_ctx* ctx = ensure_ctx();
  const int n = shim_hp();
  std::vector<double> q(3 * (size_t)n);
  check(lqro_candidates(ctx, flat(xInit1, 16).data(), flat(xInit2, 16).data(), q.data()), ctx, "lqro_candidates");
  for (int i = 0; i < n; ++i) {
    V3d p;
    p[0] = q[3 * i]; p[1] = q[3 * i + 1]; p[2] = q[3 * i + 2];
    ellipsoids.push_back(p);
  }
}

// ---- MAIN:691-706 findReachableObstacle(ellipsoids, reachablePoints, x0, x1, maxSpeed): appends survivors
template <class V3d, class State>
inline void findReachableObstacle(const std::vector<V3d>& ellipsoids, std::vector<V3d>& reachablePoints, const State& x0,
                                  const State& x1, const double& /*maxSpeed: configured at create*/) {
  lqro_ctx* ctx = ensure_ctx();
  const int n = (int)ellipsoids.size();
  std::vector<double> q(3 * (size_t)n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) q[3 * i + k] = ellipsoids[i][k];
  std::vector<int> idx(n > 0 ? n : 1);
  int cnt = 0;
  check(lqro_reachable(ctx, q.data(), n, flat(x0, 16).data(), flat(x1, 16).data(), idx.data(), &cnt), ctx, "lqro_reachable");
  for (int i = 0; i < cnt; ++i) reachablePoints.push_back(ellipsoids[idx[i]]);
}

// ---- MAIN:712-746 run_gjk(x0, x1, reachablePoints, distance, normal)
template <class V3d, class State>
inline void run_gjk(const State& x0, const State& x1, const std::vector<V3d>& reachablePoints, double& distance, V3d& normal) {
  lqro_ctx* ctx = ensure_ctx();
  const int n = (int)reachablePoints.size();
  std::vector<double> q(3 * (size_t)n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) q[3 * i + k] = reachablePoints[i][k];
  const double vrel[3] = {x0[3] - x1[3], x0[4] - x1[4], x0[5] - x1[5]};
  double sq = 0, w[3], wv[3];
  check(lqro_gjk(ctx, vrel, q.data(), n, &sq, wv, w, nullptr), ctx, "lqro_gjk");
  distance = std::sqrt(sq);                                   // MAIN:741
  normal[0] = (wv[0] - w[0]) / distance; normal[1] = (wv[1] - w[1]) / distance; normal[2] = (wv[2] - w[2]) / distance;  // MAIN:743-745
}

// ---- MAIN:749-758 pointInHull
template <class V3d, class State>
inline void pointInHull(const std::vector<V3d>& reachablePoints, bool& insideHull, const State& x0, const State& x1) {
  double distance;
  V3d normal;
  run_gjk(x0, x1, reachablePoints, distance, normal);
  insideHull = distance < 0.0001 && distance > -0.0001;
}

// ---- MAIN:809-838 convexHull(reachablePoints, x0, x1, distance, normal): no process spawn, no files
template <class V3d, class State>
inline void convexHull(const std::vector<V3d>& reachablePoints, const State& x0, const State& x1, double& distance, V3d& normal) {
  lqro_ctx* ctx = ensure_ctx();
  const int n = (int)reachablePoints.size();
  std::vector<double> q(3 * (size_t)n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) q[3 * i + k] = reachablePoints[i][k];
  const double vrel[3] = {x0[3] - x1[3], x0[4] - x1[4], x0[5] - x1[5]};
  double nrm[3];
  check(lqro_hull_escape(ctx, q.data(), n, vrel, &distance, nrm, nullptr, nullptr), ctx, "lqro_hull_escape");
  normal[0] = nrm[0]; normal[1] = nrm[1]; normal[2] = nrm[2];
}

// ---- MAIN:1085-1098 createHalfPlanes: pure narrowing arithmetic, kept on the host (FP64 -> FP32)
template <class State, class V3d, class PlaneT>
inline void createHalfPlanes(const State& x, double& distance, V3d& normal, const bool& insideHull, std::vector<PlaneT>& orcaPlanes_) {
  PlaneT p;
  const double mult = insideHull ? 1.0 : -1.0;
  p.normal = decltype(p.normal)(normal[0], normal[1], normal[2]);
  p.point = decltype(p.point)(x[3] + mult * distance * normal[0], x[4] + mult * distance * normal[1], x[5] + mult * distance * normal[2]);
  orcaPlanes_.push_back(p);
}

// ---- MAIN:1101-1112 calculateNewV(orcaPlanes_, goalVel, newVel): clears orcaPlanes_
template <class PlaneT, class V3d>
inline void calculateNewV(std::vector<PlaneT>& orcaPlanes_, const V3d& goalVel, V3d& newVel) {
  lqro_ctx* ctx = ensure_ctx();
  const int n = (int)orcaPlanes_.size();
  std::vector<float> pl(6 * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) {
    pl[6 * i + 0] = orcaPlanes_[i].point[0]; pl[6 * i + 1] = orcaPlanes_[i].point[1]; pl[6 * i + 2] = orcaPlanes_[i].point[2];
    pl[6 * i + 3] = orcaPlanes_[i].normal[0]; pl[6 * i + 4] = orcaPlanes_[i].normal[1]; pl[6 * i + 5] = orcaPlanes_[i].normal[2];
  }
  const double g[3] = {goalVel[0], goalVel[1], goalVel[2]};
  double v[3];
  check(lqro_new_velocity(ctx, 1, n > 0 ? n : 1, pl.data(), &n, g, v), ctx, "lqro_new_velocity");
  newVel[0] = v[0]; newVel[1] = v[1]; newVel[2] = v[2];
  orcaPlanes_.clear();
}

// ---- the whole avoidance block of the time loop, MAIN:1373-1407, in one device step ----------------
// Usage in the reference's _tmain:
//     lqro::Swarm swarm(NUM_QUADS, NUM_QUADS - 1);   swarm.setDynamics(A, B, qlist[0]->L, qlist[0]->E);
//     ... each time step, instead of the `for i / for j` block:
//     swarm.step(qlist);      // reads qlist[i]->x, qlist[i]->vGoal; writes qlist[i]->newV
class Swarm {
 public:
  Swarm(int n_agents, int max_neighbors, int device = 0) : n_(n_agents), k_(max_neighbors) {
    lqro_config cfg;
    lqro_default_config(&cfg);
#ifdef OBSTACLE_STEPS
    cfg.horizon = OBSTACLE_STEPS;
#endif
#ifdef NUM_POINTS
    cfg.n_samples = NUM_POINTS;
#endif
    cfg.device = device;
    cfg.n_agents = n_agents;
    cfg.max_neighbors = max_neighbors;
    check(lqro_create(&cfg, &ctx_), nullptr, "lqro_create");
    x_.resize(16 * (size_t)n_); vpref_.resize(3 * (size_t)n_); newv_.resize(3 * (size_t)n_);
    nbr_.resize((size_t)n_ * k_);
    for (int i = 0; i < n_; ++i) {  // the reference's neighbour set: every other agent in index order (MAIN:1376-1380)
      int c = 0;
      for (int j = 0; j < n_ && c < k_; ++j) if (j != i) nbr_[(size_t)i * k_ + c++] = j;
      for (; c < k_; ++c) nbr_[(size_t)i * k_ + c] = -1;
    }
  }
  ~Swarm() { lqro_destroy(ctx_); }
  Swarm(const Swarm&) = delete;
  Swarm& operator=(const Swarm&) = delete;

  template <class M16x16, class M16x4, class M4x16, class M4x3>
  void setDynamics(const M16x16& A, const M16x4& B, const M4x16& L, const M4x3& E) {
    check(lqro_set_dynamics(ctx_, flat(A, 256).data(), flat(B, 64).data(), flat(L, 64).data(), flat(E, 12).data()), ctx_, "lqro_set_dynamics");
  }
  void setNeighbors(const std::vector<int>& nbr) { nbr_ = nbr; }

  // QuadPtrVec: std::vector<Quadrotor*> with members x (State), vGoal, newV (Matrix<3,1>)
  template <class QuadPtrVec>
  void step(QuadPtrVec& qlist) {
    for (int i = 0; i < n_; ++i) {
      for (int k = 0; k < 16; ++k) x_[16 * (size_t)i + k] = qlist[i]->x[k];
      for (int k = 0; k < 3; ++k) vpref_[3 * (size_t)i + k] = qlist[i]->vGoal[k];
    }
    check(lqro_step(ctx_, x_.data(), vpref_.data(), nbr_.data(), newv_.data()), ctx_, "lqro_step");
    for (int i = 0; i < n_; ++i) for (int k = 0; k < 3; ++k) qlist[i]->newV[k] = newv_[3 * (size_t)i + k];
  }
  // ---- the second loop of the time step, MAIN:1408-1417, in one device call (lqro_sim_update) ----------------
  // Fill simModel() from the reference's globals once after setup()/controlMatrices, e.g.
  //     lqro_sim_model& m = swarm.simModel();       // pre-filled with setup()'s constants (MAIN:176-198)
  //     lqro::copy(m.L, qlist[0]->L, 64); lqro::copy(m.E, qlist[0]->E, 12); lqro::copy(m.l, qlist[0]->l, 4);
  //     lqro::copy(m.Lh, qlist[0]->Lh, 48); lqro::copy(m.Eh, qlist[0]->Eh, 9); lqro::copy(m.M, M, 256); lqro::copy(m.N, N, 36);
  //     swarm.commitSimModel();
  // then, each time step, instead of the loop MAIN:1408-1417:
  //     swarm.update(qlist, normal);                // `normal` = the reference's own normal(), MAIN:325-336
  lqro_sim_model& simModel() {
    if (!model_init_) { lqro_sim_default_model(&model_); model_init_ = true; }
    return model_;
  }
  void commitSimModel() { check(lqro_sim_set_model(ctx_, &simModel()), ctx_, "lqro_sim_set_model"); }

  // QuadPtrVec: std::vector<Quadrotor*> with members x, xTrue (State), Rot, RotTrue (Rotation), P (Matrix<16,16>), newV, pGoal,
  // vGoal (Matrix<3,1>) -- the reference's Quadrotor (MAIN:79-99).  `normal` is called 22 times per quad in the reference's
  // order (16 for propagate's process noise MAIN:450, then 6 for the observation MAIN:1413), so the CRT rand() stream is
  // consumed exactly as by the loop it replaces.  Per quad: vGoal = newV; u = findU(); propagateU(u); kalmanFilter1;
  // z = sampleGaussian(h(xTrue, RotTrue), N); kalmanFilter2; vGoal = findVGoal().  (visualize() stays with the caller.)
  template <class QuadPtrVec, class NormalFn>
  void update(QuadPtrVec& qlist, NormalFn normal) {
    const size_t n = (size_t)n_;
    sx_.resize(16 * n); sxt_.resize(16 * n); sr_.resize(9 * n); srt_.resize(9 * n); sp_.resize(256 * n);
    snv_.resize(3 * n); spg_.resize(3 * n); snr_.resize(22 * n); svg_.resize(3 * n);
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 16; ++k) { sx_[16 * i + k] = qlist[i]->x[k]; sxt_[16 * i + k] = qlist[i]->xTrue[k]; }
      for (int k = 0; k < 9; ++k) { sr_[9 * i + k] = qlist[i]->Rot[k]; srt_[9 * i + k] = qlist[i]->RotTrue[k]; }
      for (int k = 0; k < 256; ++k) sp_[256 * i + k] = qlist[i]->P[k];
      for (int k = 0; k < 3; ++k) { snv_[3 * i + k] = qlist[i]->newV[k]; spg_[3 * i + k] = qlist[i]->pGoal[k]; }
      for (int k = 0; k < 22; ++k) snr_[22 * i + k] = normal();
    }
    check(lqro_sim_update(ctx_, n_, sx_.data(), sr_.data(), sp_.data(), sxt_.data(), srt_.data(), snv_.data(), spg_.data(), snr_.data(),
                          svg_.data(), nullptr), ctx_, "lqro_sim_update");
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 16; ++k) { qlist[i]->x[k] = sx_[16 * i + k]; qlist[i]->xTrue[k] = sxt_[16 * i + k]; }
      for (int k = 0; k < 9; ++k) { qlist[i]->Rot[k] = sr_[9 * i + k]; qlist[i]->RotTrue[k] = srt_[9 * i + k]; }
      for (int k = 0; k < 256; ++k) qlist[i]->P[k] = sp_[256 * i + k];
      for (int k = 0; k < 3; ++k) qlist[i]->vGoal[k] = svg_[3 * i + k];
    }
  }
  lqro_ctx* context() { return ctx_; }

 private:
  int n_, k_;
  lqro_ctx* ctx_ = nullptr;
  std::vector<double> x_, vpref_, newv_;
  std::vector<int> nbr_;
  lqro_sim_model model_;
  bool model_init_ = false;
  std::vector<double> sx_, sxt_, sr_, srt_, sp_, snv_, spg_, snr_, svg_;
};

// ---- RVO::Agent (AGENT:84-180, RVO2-3D: the file the LP in MAIN was taken from; not in the reference's build).  The two
// members north_star names, with the reference's signatures: update() (AGENT:176-180: velocity_ = newVelocity_;
// position_ += velocity_ * timeStep) on the device through lqro_agent_update, and computeNewVelocity() (AGENT:84-151) for the
// part that is on the hot path here -- the linear programs over orcaPlanes_ (AGENT:147-150 = calculateNewV, MAIN:1101-1112);
// in this code base the planes come from the LQR-obstacles (createHalfPlanes), not from RVO's velocity obstacles.
namespace RVO {
template <class Vec3, class PlaneT>
class Agent {
 public:
  Vec3 position_, velocity_, newVelocity_, prefVelocity_;
  std::vector<PlaneT> orcaPlanes_;
  float maxSpeed_ = 100.0f;     // MAIN:1102
  float timeStep_ = 1.0f / 30;  // sim_->timeStep_
  void computeNewVelocity() {
    lqro_ctx* ctx = ensure_ctx();
    const int n = (int)orcaPlanes_.size();
    std::vector<float> pl(6 * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) { pl[6 * i + k] = orcaPlanes_[i].point[k]; pl[6 * i + 3 + k] = orcaPlanes_[i].normal[k]; }
    const double g[3] = {prefVelocity_[0], prefVelocity_[1], prefVelocity_[2]};
    double v[3];
    check(lqro_new_velocity(ctx, 1, n > 0 ? n : 1, pl.data(), &n, g, v), ctx, "lqro_new_velocity");
    newVelocity_ = Vec3(v[0], v[1], v[2]);
  }
  void update() {
    lqro_ctx* ctx = ensure_ctx();
    float p[3] = {position_[0], position_[1], position_[2]}, v[3] = {velocity_[0], velocity_[1], velocity_[2]};
    const float nv[3] = {newVelocity_[0], newVelocity_[1], newVelocity_[2]};
    check(lqro_agent_update(ctx, 1, p, v, nv, timeStep_), ctx, "lqro_agent_update");
    position_ = Vec3(p[0], p[1], p[2]);
    velocity_ = Vec3(v[0], v[1], v[2]);
  }
};
}  // namespace RVO

// element-wise copy out of a reference Matrix (operator[] const returns by value, MAT:58-61)
template <class M>
inline void copy(double* dst, const M& m, int n) { for (int i = 0; i < n; ++i) dst[i] = m[i]; }

}  // namespace lqro
