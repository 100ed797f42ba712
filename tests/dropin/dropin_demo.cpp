// tests/dropin/dropin_demo.cpp -- the reference's hot loop (MAIN:1371-1407) written against include/lqro_dropin.hpp.
// Built two ways by tests/test_dropin.py:
//   * -DUSE_REFERENCE_HEADERS with -I/root/reference/{include,QuadrotorHoverController}: the shims are instantiated
//     with the reference's OWN Matrix<R,C>, Vector3 and Plane types (build container only);
//   * stand-alone with the minimal stand-ins below (same interface: row-major storage, operator[] / operator()),
//     which is what travels to the GPU box, where the program runs one step per scenario state read from stdin
//     and prints newV for the test to compare with the golden vectors.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef USE_REFERENCE_HEADERS
#include "matrix.h"
#include "Vector3.h"
#else
template <size_t R, size_t C = 1>
class Matrix {
  double e_[R * C];
 public:
  double& operator[](size_t i) { return e_[i]; }
  double operator[](size_t i) const { return e_[i]; }
  double& operator()(size_t r, size_t c) { return e_[r * C + c]; }
  double operator()(size_t r, size_t c) const { return e_[r * C + c]; }
};
class Vector3 {
  float v_[3];
 public:
  Vector3() { v_[0] = v_[1] = v_[2] = 0.f; }
  Vector3(double x, double y, double z) { v_[0] = (float)x; v_[1] = (float)y; v_[2] = (float)z; }
  float operator[](size_t i) const { return v_[i]; }
  float& operator[](size_t i) { return v_[i]; }
};
#endif
struct Plane { Vector3 point; Vector3 normal; };  // MAIN:844-855

#define OBSTACLE_STEPS 30
#define NUM_POINTS 50
#define XYRADIUS 0.26
#define ZRADIUS 0.75
#include "lqro_dropin.hpp"

typedef Matrix<16> State;
struct Quadrotor { State x, xTrue; Matrix<3, 3> Rot, RotTrue; Matrix<16, 16> P; Matrix<3, 1> vGoal, newV, pGoal; };
static double demo_normal() { return 0.0; }  // stands in for the reference's normal() (MAIN:325-336)

static bool read_doubles(double* d, int n) { for (int i = 0; i < n; ++i) if (scanf("%lf", d + i) != 1) return false; return true; }

int main(int argc, char** argv) {
  // stdin: n, then A(256) B(64) L(64) E(12), then per step: n x (x[16], vGoal[3]); stdout: newV per agent per step
  int n = 0;
  if (scanf("%d", &n) != 1 || n < 2) return 2;
  Matrix<16, 16> A; Matrix<16, 4> B; Matrix<4, 16> L; Matrix<4, 3> E;
  if (!read_doubles(&A[0], 256) || !read_doubles(&B[0], 64) || !read_doubles(&L[0], 64) || !read_doubles(&E[0], 12)) return 2;
  std::vector<Quadrotor*> qlist(n);
  for (int i = 0; i < n; ++i) qlist[i] = new Quadrotor();
  lqro::Swarm swarm(n, n - 1);
  swarm.setDynamics(A, B, L, E);
  lqro::setDynamics(A, B, L, E);  // per-function shims use their own small context
  if (argc > 1) {  // instantiates the simulator-form update shims (MAIN:1408-1417); exercised by tests/test_gpu_sim.py via the C ABI
    lqro_sim_model& m = swarm.simModel();
    lqro::copy(m.L, L, 64); lqro::copy(m.E, E, 12);
    swarm.commitSimModel();
    swarm.update(qlist, demo_normal);
    Matrix<4> uGoal, lvec; Matrix<16> xstar, cvec; Matrix<3, 16> Lh; Matrix<3, 3> Eh;
    for (int i = 0; i < 4; ++i) uGoal[i] = 9.80665 * 0.5 / 4;
    lqro::controlMatrices(uGoal, xstar, A, B, cvec, L, E, lvec, Lh, Eh);  // SIM2:43 argument list
    return 0;
  }
  while (true) {
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) ok = read_doubles(&qlist[i]->x[0], 16) && read_doubles(&qlist[i]->vGoal[0], 3);
    if (!ok) break;
    swarm.step(qlist);  // replaces the whole for i / for j block
    for (int i = 0; i < n; ++i) printf("%.17g %.17g %.17g\n", qlist[i]->newV[0], qlist[i]->newV[1], qlist[i]->newV[2]);
    // and the same for agent 0 through the per-function shims, in the reference's own call order (MAIN:1376-1406)
    std::vector<Plane> orcaPlanes_;
    std::vector<Matrix<3, 1> > ellipsoids, reachablePoints;
    bool insideHull = false; double distance = 0; Matrix<3, 1> normalVector;
    for (int j = 1; j < n; ++j) {
      ellipsoids.clear(); reachablePoints.clear();
      lqro::createObstacleAll(ellipsoids, qlist[0]->x, qlist[j]->x);
      lqro::findReachableObstacle(ellipsoids, reachablePoints, qlist[0]->x, qlist[j]->x, 30.0);
      if ((int)reachablePoints.size() > 4) {
        lqro::pointInHull(reachablePoints, insideHull, qlist[0]->x, qlist[j]->x);
        if (insideHull) lqro::convexHull(reachablePoints, qlist[0]->x, qlist[j]->x, distance, normalVector);
        else lqro::run_gjk(qlist[0]->x, qlist[j]->x, reachablePoints, distance, normalVector);
        distance *= 0.5;
        lqro::createHalfPlanes(qlist[0]->x, distance, normalVector, insideHull, orcaPlanes_);
      }
    }
    Matrix<3, 1> v;
    lqro::calculateNewV(orcaPlanes_, qlist[0]->vGoal, v);
    printf("shim %.17g %.17g %.17g\n", v[0], v[1], v[2]);
  }
  return 0;
}
