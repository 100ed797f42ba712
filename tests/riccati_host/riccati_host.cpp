// tests/riccati_host/riccati_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the KERNEL SOURCE lqr_obstacles_b200/csrc/lqro_riccati.cuh (riccati_warp: the batched Riccati solve, one warp per
// problem, recursion state in registers, shuffles for the row exchange and the cooperative m x m inverse) as plain C++ on top
// of tests/hull_host/warp_emul.h (32 host threads in lockstep stand in for the lanes), so that its logic can be checked against
// the oracle on a machine without a GPU (tests/test_riccati_host.py).  The product (liblqro.so) never links this.
#include "../hull_host/warp_emul.h"

#include "../../lqr_obstacles_b200/csrc/lqro_riccati.cuh"

#include <stdlib.h>

extern "C" int riccatihost_solve(const double* A, const double* B, double qv, double qp, double r, int iters, double* L, double* E, double* Lh,
                                 double* Eh) {
  RiccatiWs* w = (RiccatiWs*)aligned_alloc(16, sizeof(RiccatiWs));
  memset(w, 0, sizeof(RiccatiWs));
  emu_run_warp([&](int) { riccati_warp(*w, A, B, qv, qp, r, iters, L, E, Lh, Eh); });
  free(w);
  return 0;
}
