// tests/pair_host/pair_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the KERNEL SOURCE of stage 2 of the pair kernel -- lqr_obstacles_b200/csrc/lqro_pair.cuh (pair_translate_steps,
// pair_scan_cloud: Translate_k in the reference's order, candidates + reachable filter into the warp's cloud) and
// lqr_obstacles_b200/csrc/lqro_gjk.cuh (gjk_warp_lanes: the lane-parallel GJK, and gjk_warp, the scalar ablation) -- as plain C++
// on top of tests/hull_host/warp_emul.h (32 host threads in lockstep stand in for the lanes), so that the logic of what k_pairs
// runs per pair can be checked bit for bit against the oracle on a machine without a GPU (tests/test_pair_host.py).
// The product (liblqro.so) never links this.
#include "../hull_host/warp_emul.h"

static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }

#include "../../lqr_obstacles_b200/csrc/lqro_common.cuh"
#include "../../lqr_obstacles_b200/csrc/lqro_gjk.cuh"
#include "../../lqr_obstacles_b200/csrc/lqro_pair.cuh"

#include <stdlib.h>

// One pair through stage 2.  NFt [16][3H], T [H][9], sph [P][3] as the device tables (lqro_common.cuh); kmask: the surviving
// horizon steps (all of them gives the same cloud as any exact pruning); templ: 1 = the compile-time P = 50 walk (needs P == 50),
// 0 = the generic walk; cap: points of the cloud kept in "shared memory" (the rest goes to the overflow arrays, as in k_pairs);
// lanes: 1 = gjk_warp_lanes, 0 = gjk_warp.
// out: cloud [n][3], gjk [8] = d^2, witness on v_rel [3], witness on the hull [3], number of G-tests.  Returns n.
extern "C" int pairhost_stage2(int H, int P, int templ, const double* NFt, const double* T, const double* sph, const double* xi, const double* xj,
                               double max_speed, unsigned long long kmask, int cap, int lanes, double* cloud, double* gjk) {
  const int HP = H * P;
  std::vector<double> sphx(P), sphy(P), sphz(P), tr(3 * H, 0.0), dx(16), sx(cap), sy(cap), sz(cap), gx(HP), gy(HP), gz(HP), scratch(LQRO_GJK_WARP_DOUBLES, 0.0);
  std::vector<int> klist(H, 0);
  for (int s = 0; s < P; ++s) { sphx[s] = sph[3 * s]; sphy[s] = sph[3 * s + 1]; sphz[s] = sph[3 * s + 2]; }
  for (int m = 0; m < 16; ++m) dx[m] = xi[m] - xj[m];
  int nk = 0;
  for (int k = 0; k < H; ++k) if ((kmask >> k) & 1ull) klist[nk++] = k;
  PairTables t;
  t.NFt = NFt; t.T = T; t.sphx = sphx.data(); t.sphy = sphy.data(); t.sphz = sphz.data(); t.CG = nullptr; t.prn = nullptr;
  CloudView cl;
  cl.sx = sx.data(); cl.sy = sy.data(); cl.sz = sz.data(); cl.gx = gx.data(); cl.gy = gy.data(); cl.gz = gz.data(); cl.cap = cap;
  const double r2 = max_speed * max_speed, inv_r2 = 1.0 / r2, r2lo = r2 * (1.0 - 1.0e-9), r2hi = r2 * (1.0 + 1.0e-9);
  int n_out = 0;
  emu_run_warp([&](int lane) {
    pair_translate_steps(t.NFt, dx.data(), klist.data(), nk, 3 * H, tr.data());
    __syncwarp();
    const double vr0 = dx[3], vr1 = dx[4], vr2 = dx[5];
    int n;
    if (templ) n = pair_scan_cloud<50>(t, tr.data(), klist.data(), nk, P, vr0, vr1, vr2, r2, inv_r2, r2lo, r2hi, cl);
    else n = pair_scan_cloud<0>(t, tr.data(), klist.data(), nk, P, vr0, vr1, vr2, r2, inv_r2, r2lo, r2hi, cl);
    double w1[3] = {0, 0, 0}, w2[3] = {0, 0, 0};
    int it = 0;
    double sq = 0;
    if (n > 4) {  // "GJK errors with less than 4", MAIN:1393
      if (lanes) sq = gjk_warp_lanes(vr0, vr1, vr2, cl, n, scratch.data(), w1, w2, &it);
      else sq = gjk_warp(vr0, vr1, vr2, cl, n, w1, w2, &it);
    }
    if (lane == 0) {
      n_out = n;
      gjk[0] = sq;
      for (int d = 0; d < 3; ++d) { gjk[1 + d] = w1[d]; gjk[4 + d] = w2[d]; }
      gjk[7] = it;
    }
  });
  for (int m = 0; m < n_out; ++m) cl.get(m, cloud[3 * m], cloud[3 * m + 1], cloud[3 * m + 2]);
  return n_out;
}

// GJK alone on an arbitrary cloud (AoS [n][3])
extern "C" void pairhost_gjk(const double* vrel, const double* pts, int n, int cap, int lanes, double* gjk) {
  std::vector<double> sx(cap > 0 ? cap : 1), sy(cap > 0 ? cap : 1), sz(cap > 0 ? cap : 1), gx(n), gy(n), gz(n), scratch(LQRO_GJK_WARP_DOUBLES, 0.0);
  CloudView cl;
  cl.sx = sx.data(); cl.sy = sy.data(); cl.sz = sz.data(); cl.gx = gx.data(); cl.gy = gy.data(); cl.gz = gz.data(); cl.cap = cap;
  for (int m = 0; m < n; ++m) cl.put(m, pts[3 * m], pts[3 * m + 1], pts[3 * m + 2]);
  emu_run_warp([&](int lane) {
    double w1[3], w2[3];
    int it = 0;
    const double sq = lanes ? gjk_warp_lanes(vrel[0], vrel[1], vrel[2], cl, n, scratch.data(), w1, w2, &it) : gjk_warp(vrel[0], vrel[1], vrel[2], cl, n, w1, w2, &it);
    if (lane == 0) {
      gjk[0] = sq;
      for (int d = 0; d < 3; ++d) { gjk[1 + d] = w1[d]; gjk[4 + d] = w2[d]; }
      gjk[7] = it;
    }
  });
}
