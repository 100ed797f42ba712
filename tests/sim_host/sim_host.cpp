// tests/sim_host/sim_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the kernel source lqr_obstacles_b200/csrc/lqro_sim.cuh as plain C++ (its warp policy collapses to one
// sequential lane) so that the LOGIC of the simulator-form update kernel can be checked against the compiled
// reference on a machine without a GPU (tests/test_sim_host.py).  The product (liblqro.so) never links this;
// there is no CPU fallback.
#include "../../lqr_obstacles_b200/csrc/lqro_sim.cuh"

static int g_use_m_diag = 1;  // tests switch the diagonal-M shortcut of sim_flow off to compare both forms
extern "C" void simhost_use_m_diag(int on) { g_use_m_diag = on; }
static int m_is_diag(const LqroSimModel* md) {
  if (!g_use_m_diag) return 0;
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) if (i != j && md->M[i * 16 + j] != 0.0) return 0;
  return 1;
}
extern "C" int simhost_model_doubles(void) { return (int)(sizeof(LqroSimModel) / sizeof(double)); }

extern "C" void simhost_update(int n, const LqroSimModel* md, double* x, double* Rot, double* P, double* xTrue, double* RotTrue,
                               const double* newV, const double* pGoal, const double* normals, double* vGoal, double* u) {
  static LqroSimWsA wa; static LqroSimWsB wb; static LqroSimWsC wc;
  for (int a = 0; a < n; ++a) {
    double MM[2][256], v1[2][16], Ae[256], uu[4], vG[3];
    // control (k_sim_control)
    sim_find_u(uu, x + 16 * a, Rot + 9 * a, newV + 3 * a, *md);
    // flow, task 0 = true state, task 1 = estimate (k_sim_flow)
    for (int est = 0; est < 2; ++est) {
      for (int e = 0; e < 16; ++e) wa.xs[e] = est ? x[16 * a + e] : xTrue[16 * a + e];
      for (int e = 0; e < 9; ++e) wa.Rs[e] = est ? Rot[9 * a + e] : RotTrue[9 * a + e];
      for (int e = 0; e < 4; ++e) wa.u[e] = uu[e];
      wa.m_diag = m_is_diag(md);
      sim_phase_flow<SimWarp>(wa, MM[est], v1[est], est ? Ae : nullptr, *md);
    }
    // noise (k_sim_noise)
    for (int e = 0; e < 256; ++e) wb.D[e] = MM[0][e];
    for (int e = 0; e < 16; ++e) wb.nrm[e] = normals[22 * a + e];
    sim_phase_b<SimWarp>(wb);
    // filter (k_sim_filter)
    for (int e = 0; e < 256; ++e) wc.P[e] = P[256 * a + e];
    for (int e = 0; e < 16; ++e) { wc.x[e] = x[16 * a + e]; wc.xT[e] = xTrue[16 * a + e]; }
    for (int e = 0; e < 9; ++e) { wc.R[e] = Rot[9 * a + e]; wc.RT[e] = RotTrue[9 * a + e]; }
    for (int e = 0; e < 6; ++e) wc.nrm[e] = normals[22 * a + 16 + e];
    sim_phase_filter<SimWarp>(wc, Ae, MM[1], v1[1], v1[0], wb.samp, pGoal + 3 * a, vG, *md);
    for (int e = 0; e < 256; ++e) P[256 * a + e] = wc.P[e];
    for (int e = 0; e < 16; ++e) { x[16 * a + e] = wc.x[e]; xTrue[16 * a + e] = wc.xT[e]; }
    for (int e = 0; e < 9; ++e) { Rot[9 * a + e] = wc.R[e]; RotTrue[9 * a + e] = wc.RT[e]; }
    for (int e = 0; e < 3; ++e) vGoal[3 * a + e] = vG[e];
    if (u) for (int e = 0; e < 4; ++e) u[4 * a + e] = uu[e];
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
extern "C" void simhost_linearize(const LqroSimModel* md, const double* x, const double* Rot, const double* u, double* A, double* B, double* c) {
  static LqroSimWsA w;
  for (int e = 0; e < 16; ++e) w.xs[e] = x[e];
  for (int e = 0; e < 9; ++e) w.Rs[e] = Rot[e];
  for (int e = 0; e < 4; ++e) w.u[e] = u[e];
  w.m_diag = m_is_diag(md);
  sim_phase_linearize<SimWarp>(w, A, B, c, *md);
}
extern "C" void simhost_init(int n, const LqroSimModel* md, const double* x, const double* Rot, const double* P, const double* pGoal,
                             const double* normals16, double* xTrue, double* RotTrue, double* vGoal) {
  static LqroSimWsB w;
  for (int a = 0; a < n; ++a) {
    for (int e = 0; e < 256; ++e) w.D[e] = P[256 * a + e];
    for (int e = 0; e < 16; ++e) { w.nrm[e] = normals16[16 * a + e]; w.zero[e] = x[16 * a + e]; }
    sim_phase_init<SimWarp>(w, x + 16 * a, Rot + 9 * a, pGoal + 3 * a, xTrue + 16 * a, RotTrue + 9 * a, vGoal + 3 * a, *md);
  }
}
// exp(q) and exp(q*0.5) through the shared-Pade routine; returns 1 when the shared path was taken, else 0 (then the
// outputs come from two separate sim_exp calls, as sim_flow falls back)
extern "C" int simhost_exp_with_half(const double* q, double* full, double* half) {
  static double ws[4 * 256], a[256]; int perm[32]; double small[16];
  for (int e = 0; e < 256; ++e) a[e] = q[e];
  if (sim_exp_with_half<SimSeq, 16>(full, half, a, ws, perm, small)) return 1;
  sim_exp<SimSeq, 16>(full, a, ws, perm, small);
  for (int e = 0; e < 256; ++e) a[e] = q[e] * 0.5;
  sim_exp<SimSeq, 16>(half, a, ws, perm, small);
  return 0;
}
extern "C" void simhost_force_general_exp(int on) { sim_force_general_exp = on; }
