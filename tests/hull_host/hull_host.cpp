// tests/hull_host/hull_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the KERNEL SOURCE lqr_obstacles_b200/csrc/lqro_hull_core.cuh as plain C++ on top of warp_emul.h (32 host
// threads in lockstep stand in for the 32 lanes of the hull warp) so that the logic of the CUDA hull -- the pending
// ring, the wave-parallel horizon search, the match.any ridge matching, the bucket-leader appends -- can be checked
// facet for facet against oracle/hull_cpu.cpp on a machine without a GPU (tests/test_hull_host.py).
// The product (liblqro.so) never links this; there is no CPU fallback.
#include "warp_emul.h"

#include "../../lqr_obstacles_b200/csrc/lqro_hull_core.cuh"

#include <stdlib.h>

// Hull of an AoS cloud.  mode bit 0: apply the 6-digit rendering first (MAIN:765-767); bit 1: two warps per hull (a main and a
// helper warp, as k_hull runs it: planes + ridge matching and the speculative horizon search on the helper).  cap_f <= 0: full-size pool.
// ring > 0: pending ring of that many entries (power of two).
// Outputs facets in qhull order (planes nf x 4, verts nf x 3 as point indices), info = {nf, degenerate, fail,
// ninserted, best facet position}, and when vrel != NULL the nearest-facet result (escape[0] distance, [1..3] normal).
extern "C" int hullhost_build(const double* pts, int n, int mode, int cap_f, int ring, double* planes, int* verts, int maxf, int* info5,
                              const double* vrel, double* escape4) {
  const int cap_pts = n > 4 ? n : 4;
  const int cf = cap_f > 0 ? cap_f : hull_full_facets(cap_pts);
  char* coords = (char*)calloc(1, hull_align16(3 * sizeof(double) * (size_t)cap_pts) + 64);
  char* hot = (char*)calloc(1, hull_hot_bytes(cap_pts, cf) + 64);
  char* cold = (char*)calloc(1, hull_cold_bytes(cap_pts, cf) + 64);
  HullCtl ctl;
  memset(&ctl, 0, sizeof ctl);
  HullMail mail;
  memset(&mail, 0, sizeof mail);
  const bool coop = (mode & 2) != 0;
  emu_run_block(coop ? 2 : 1, [&](int warp, int lane) {
    HullWs w;
    hull_carve_coords(coords, cap_pts, w);
    hull_carve_hot(hot, cap_pts, cf, w);
    hull_carve_cold(cold, cap_pts, cf, w);
    if (ring > 0 && ring - 1 < w.ring_mask) w.ring_mask = ring - 1;  // tests: a smaller pending ring (power of two)
    for (int m = lane; m < n; m += 32) {
      const double a = pts[3 * m], b = pts[3 * m + 1], c = pts[3 * m + 2];
      w.ux[m] = a; w.uy[m] = b; w.uz[m] = c;
      if (mode & 1) { w.hx[m] = round6_dev(a); w.hy[m] = round6_dev(b); w.hz[m] = round6_dev(c); }
      else { w.hx[m] = a; w.hy[m] = b; w.hz[m] = c; }
    }
    if (coop && warp == 1) {   // the helper warp: waits until the main warp has set up the cloud, then serves it
      emu_block_barrier();
      hull_helper(w, ctl, mail);
      return;
    }
    __syncwarp();
    if (coop) { emu_block_barrier(); hull_build_t<true>(w, ctl, &mail, n, true); }
    else hull_build(w, ctl, n, true);
    if (lane == 0) { info5[0] = ctl.fail ? 0 : ctl.nfacets_out; info5[1] = ctl.degenerate; info5[2] = ctl.fail; info5[3] = ctl.ninserted; info5[4] = -1; }
    if (ctl.fail) return;
    for (int q = lane; q < ctl.nfacets_out && q < maxf; q += 32) {
      const int f = w.order[q];
      const HPlane F = w.pl[f];
      const HTopo T = w.tp[f];
      if (planes) { planes[4 * q] = F.n0; planes[4 * q + 1] = F.n1; planes[4 * q + 2] = F.n2; planes[4 * q + 3] = F.off; }
      if (verts) { verts[3 * q] = w.vpoint[T.v[0]]; verts[3 * q + 1] = w.vpoint[T.v[1]]; verts[3 * q + 2] = w.vpoint[T.v[2]]; }
    }
    if (vrel) {
      double distance, normal[3];
      const int bf = hull_nearest_facet(w, ctl, w.ux, w.uy, w.uz, vrel[0], vrel[1], vrel[2], &distance, normal);
      const int pos = hull_list_position(w, ctl, bf);
      if (lane == 0) { escape4[0] = distance; escape4[1] = normal[0]; escape4[2] = normal[1]; escape4[3] = normal[2]; info5[4] = pos; }
    }
  });
  free(coords); free(hot); free(cold);
  return 0;
}

extern "C" void hullhost_coverage(long long* out16, int reset) { for (int i = 0; i < 16; ++i) { out16[i] = g_hull_cover[i]; if (reset) g_hull_cover[i] = 0; } }

extern "C" double hullhost_round6(double x) { return round6_dev(x); }
