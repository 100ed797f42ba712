// tests/dropin/ref_tmain_dropin.cpp -- TEST INFRASTRUCTURE ONLY (built only where /root/reference exists).
// The reference's OWN `_tmain` -- scenario set-up, the time loop MAIN:1371-1418, shutdown -- compiled UNMODIFIED against
// include/lqro_dropin.hpp: every hot-path call of the loop (findFG, createObstacle, findReachableObstacle, pointInHull,
// run_gjk, convexHull, createHalfPlanes, calculateNewV; createSpheres before it) binds to the shim with the reference's
// own signature and runs on the GPU through the C ABI; everything else (Quadrotor, controlMatrices, propagate, the Kalman
// filter, the noise stream) is the reference's CPU code.
//
// How: the reference translation unit is #included where it lies under /root/reference with its nine hot-path functions
// (and its _tmain) renamed by macros, so that it supplies the types, globals and the update code; then the text of _tmain
// (tests/dropin/_gen/tmain.inc, extracted verbatim at build time by extract_tmain.py, git-ignored) is compiled with the
// original names bound to lqro::.  The Callisto stub records the key frames; main() prints them for tests/test_dropin.py,
// which compares the trajectory with the one the untouched reference produces (tests/golden/tmain_keyframes.npy).
#define findFG ref_cpu_findFG
#define createSpheres ref_cpu_createSpheres
#define createObstacle ref_cpu_createObstacle
#define findReachableObstacle ref_cpu_findReachableObstacle
#define pointInHull ref_cpu_pointInHull
#define run_gjk ref_cpu_run_gjk
#define convexHull ref_cpu_convexHull
#define createHalfPlanes ref_cpu_createHalfPlanes
#define calculateNewV ref_cpu_calculateNewV
#include "LQRObstaclesPositionControl(Working).cpp"
#undef findFG
#undef createSpheres
#undef createObstacle
#undef findReachableObstacle
#undef pointInHull
#undef run_gjk
#undef convexHull
#undef createHalfPlanes
#undef calculateNewV

#include <cstdio>
#include "lqro_dropin.hpp"
using lqro::findFG;
using lqro::createSpheres;
using lqro::createObstacle;
using lqro::findReachableObstacle;
using lqro::pointInHull;
using lqro::run_gjk;
using lqro::convexHull;
using lqro::createHalfPlanes;
using lqro::calculateNewV;

#undef _tmain
#define _tmain lqro_dropin_tmain
#include "_gen/tmain.inc"

// what oracle/shims/pre.h expects from the harness: the MSVC CRT generator and the qconvex route (unused: convexHull is the shim)
static unsigned int g_seed = 1u;
extern "C" int lqr_msvc_rand(void) { g_seed = g_seed * 214013u + 2531011u; return (int)((g_seed >> 16) & 0x7fffu); }
extern "C" void lqr_msvc_srand(unsigned int s) { g_seed = s; }
extern "C" int lqr_ref_system(const char*) { return -1; }
extern "C" int lqr_ref_num_keyframes(void);
extern "C" void lqr_ref_get_keyframe(int i, int* group, float* t, float* p);

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: ref_tmain_dropin <keyframes.txt>   (stdin from /dev/null)\n"); return 2; }
  lqr_msvc_srand(1u);
  const int rc = lqro_dropin_tmain(0, 0);
  FILE* f = std::fopen(argv[1], "w");
  if (!f) return 3;
  const int n = lqr_ref_num_keyframes();
  for (int i = 0; i < n; ++i) {
    int g; float t, p[3];
    lqr_ref_get_keyframe(i, &g, &t, p);
    std::fprintf(f, "%d %.9g %.9g %.9g %.9g\n", g, t, p[0], p[1], p[2]);
  }
  std::fclose(f);
  return rc;
}
