// tests/sim_host/sim_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the kernel source lqr_obstacles_b200/csrc/lqro_sim.cuh as plain C++ (its warp policy collapses to one
// sequential lane) so that the LOGIC of the simulator-form update kernel can be checked against the compiled
// reference on a machine without a GPU (tests/test_sim_host.py).  The product (liblqro.so) never links this;
// there is no CPU fallback.
#include "../../lqr_obstacles_b200/csrc/lqro_sim.cuh"

extern "C" int simhost_model_doubles(void) { return (int)(sizeof(LqroSimModel) / sizeof(double)); }

extern "C" void simhost_update(int n, const LqroSimModel* md, double* x, double* Rot, double* P, double* xTrue, double* RotTrue,
                               const double* newV, const double* pGoal, const double* normals, double* vGoal, double* u) {
  static LqroSimWs w;
  for (int a = 0; a < n; ++a) {
    for (int e = 0; e < 256; ++e) w.P[e] = P[256 * a + e];
    for (int e = 0; e < 16; ++e) { w.x[e] = x[16 * a + e]; w.xT[e] = xTrue[16 * a + e]; }
    for (int e = 0; e < 9; ++e) { w.R[e] = Rot[9 * a + e]; w.RT[e] = RotTrue[9 * a + e]; }
    for (int e = 0; e < 22; ++e) w.nrm[e] = normals[22 * a + e];
    double vG[3];
    sim_update_agent<SimWarp>(w, newV + 3 * a, pGoal + 3 * a, vG, *md);
    for (int e = 0; e < 256; ++e) P[256 * a + e] = w.P[e];
    for (int e = 0; e < 16; ++e) { x[16 * a + e] = w.x[e]; xTrue[16 * a + e] = w.xT[e]; }
    for (int e = 0; e < 9; ++e) { Rot[9 * a + e] = w.R[e]; RotTrue[9 * a + e] = w.RT[e]; }
    for (int e = 0; e < 3; ++e) vGoal[3 * a + e] = vG[e];
    if (u) for (int e = 0; e < 4; ++e) u[4 * a + e] = w.u[e];
  }
}
extern "C" void simhost_exp16(const double* q, double* out) {
  static double ws[4 * 256], a[256]; int perm[32]; double small[16];
  for (int e = 0; e < 256; ++e) a[e] = q[e];  // sim_exp scales its argument in place
  sim_exp<SimSeq, 16>(out, a, ws, perm, small);
}
extern "C" void simhost_exp3(const double* q, double* out) { sim_exp3(out, q); }
extern "C" void simhost_jacobi16(const double* q, double* V, double* D) {
  for (int e = 0; e < 256; ++e) D[e] = q[e];
  sim_jacobi<SimSeq, 16>(D, V);
}
extern "C" void simhost_f(const LqroSimModel* md, const double* x, const double* Rot, const double* u, double* xdot) {
  sim_f(xdot, x, Rot, u, *md);
}
