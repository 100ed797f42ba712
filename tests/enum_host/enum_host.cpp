// tests/enum_host/enum_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the pair enumeration of the pair kernel (lqr_obstacles_b200/csrc/lqro_enum.cuh: plain integer code, host + device) with
// g++ so that tests/test_enum_host.py can walk it on a machine without a GPU.  The product (liblqro.so) never links this.
#include "../hull_host/warp_emul.h"
#include "../../lqr_obstacles_b200/csrc/lqro_enum.cuh"

// Walks the whole enumeration of n_local agents x K slots (first chunk `first` slots wide, agents split at `split`, 0 = one
// segment).  out_il / out_kk [total]: the pair of every index, -1 / -1 for the padding; returns total (padded to 32-pair batches).
extern "C" long long enumhost_walk(int n_local, int K, int first, int split, int* out_il, int* out_kk, long long cap) {
  LqroParams p{};
  p.n_local = n_local; p.K = K; p.first_slots = first; p.seg_split = split;
  const long long total_a = split > 0 ? slot_major_total(split, K, first) : 0;
  const long long total = total_a + slot_major_total(n_local - (split > 0 ? split : 0), K, first);
  for (long long idx = 0; idx < total && idx < cap; ++idx) {
    int il = -1, kk = -1;
    if (!slot_major_pair<long long>(p, total_a, idx, il, kk)) il = kk = -1;
    out_il[idx] = il; out_kk[idx] = kk;
  }
  return total;
}
extern "C" long long enumhost_prefix(int n_local, int K, int first, int slots) { return slot_major_prefix(n_local, K, first, slots); }
// The batches of a launch: base / width of every batch up to the first one past the end (and `extra` more, which must all be
// rejected).  Returns the number of accepted batches, -1 if a batch after a rejected one is accepted.
extern "C" int enumhost_batches(int n_local, int K, int first, int split, int narrow_slots, int width, int* out_base, int* out_bw, int cap, int extra) {
  const BatchPlan b = make_batch_plan(n_local, K, first, split, narrow_slots, width);
  int n = 0;
  for (int batch = 0; batch < cap; ++batch) {
    int base = -1, bw = -1;
    if (!slot_major_batch(b, batch, base, bw)) {
      for (int e = 1; e <= extra; ++e) if (slot_major_batch(b, batch + e * 977, base, bw)) return -1;
      if (slot_major_batch(b, 0x7fffffff, base, bw)) return -1;
      return n;
    }
    out_base[n] = base; out_bw[n] = bw; ++n;
  }
  return n;
}

