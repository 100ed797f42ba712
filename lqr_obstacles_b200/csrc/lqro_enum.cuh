// lqro_enum.cuh -- the order in which a pair-kernel launch walks the (agent, neighbour slot) pairs of a rank (used by k_pairs in
// lqro_step.cuh and by the host when it sizes the batches, lqro_api.cu).  Plain integer arithmetic, host + device: the same source
// is checked on the CPU (tests/test_enum_host.py: every pair exactly once, batches inside one segment, padding rejected).
// The reference walks agent after agent, neighbour after neighbour (MAIN:1376-1379); the order has no influence on any result.
#pragma once
#include "lqro_common.cuh"

// slot-major enumeration: the first chunk covers the p.first_slots nearest neighbour slots of every agent, the following ones
// double in width up to 32 slots (first_slots = 1: slots 0 | 1 | 2-3 | 4-7 | 8-15 | 16-31 | 32-63); within a chunk pairs are
// agent-major.  Neighbour lists are sorted by distance, so the inside pairs -- the ones that need a hull, a ~1 ms sequential
// job each -- sit in the first chunks and reach the hull queue in the first tenth of the pair kernel, where the polling hull
// launch picks them up (launch_step).
// idx-th pair of the whole enumeration -> (local agent, slot); returns false past the end (chunks start on a batch boundary).
__host__ __device__ __forceinline__ int slot_chunk_width(int K, int c0, int first) {
  const int w = (c0 == 0) ? first : (c0 < 32 ? c0 : 32);
  return (K - c0 < w) ? (K - c0) : w;
}
__host__ __device__ __forceinline__ int slot_major_log2(int w) {
#ifdef __CUDA_ARCH__
  return 31 - __clz(w);
#else
  int l = 0;
  while ((1 << (l + 1)) <= w) ++l;
  return l;
#endif
}
// I: long long, or int when the padded pair count of the launch stays below 2^31 (the kernel: lqro_create checks)
template <typename I>
__host__ __device__ __forceinline__ bool slot_major_pair_n(int n_agents, int K, int first, I idx, int& il, int& kk) {
  int c0 = 0;
  while (c0 < K) {
    const int wdt = slot_chunk_width(K, c0, first);
    const I cnt = (I)n_agents * wdt;
    const I padded = (cnt + 31) & ~(I)31;  // chunks start on a batch boundary
    if (idx < padded) {
      if (idx >= cnt) return false;
      // chunk widths are powers of two except possibly the first: a shift instead of the 64-bit division routine
      if ((wdt & (wdt - 1)) == 0) il = (int)(idx >> slot_major_log2(wdt));
      else il = idx < (I)0x7fffffff ? (int)((unsigned)idx / (unsigned)wdt) : (int)(idx / wdt);
      kk = c0 + (int)(idx - (I)il * wdt);
      return true;
    }
    idx -= padded;
    c0 += wdt;
  }
  return false;
}
// Two segments of agents, [0, split) and [split, n_local), each enumerated slot-major, the first one first (split = 0: one segment).
// Used when the neighbour table of a step is still on its way to the device while the pair kernel runs (StepSync::nbr_mark):
// the rows of the short first segment arrive first and are worked on while the rest is uploaded.  total_a = padded pairs of the
// first segment (slot_major_total(split, ...)).
template <typename I>
__host__ __device__ __forceinline__ bool slot_major_pair(const LqroParams& p, I total_a, I idx, int& il, int& kk) {
  if (p.seg_split <= 0) return slot_major_pair_n(p.n_local, p.K, p.first_slots, idx, il, kk);
  if (idx < total_a) return slot_major_pair_n(p.seg_split, p.K, p.first_slots, idx, il, kk);
  if (!slot_major_pair_n(p.n_local - p.seg_split, p.K, p.first_slots, idx - total_a, il, kk)) return false;
  il += p.seg_split;
  return true;
}
__host__ __device__ inline long long slot_major_total(int n_local, int K, int first) {
  long long t = 0;
  for (int c0 = 0; c0 < K; c0 += slot_chunk_width(K, c0, first)) t += (((long long)n_local * slot_chunk_width(K, c0, first)) + 31) & ~31ll;
  return t;
}
// pairs (padded to batches) of the chunks that start below slot `slots`
__host__ __device__ inline long long slot_major_prefix(int n_local, int K, int first, int slots) {
  long long t = 0;
  for (int c0 = 0; c0 < K && c0 < slots; c0 += slot_chunk_width(K, c0, first)) t += (((long long)n_local * slot_chunk_width(K, c0, first)) + 31) & ~31ll;
  return t;
}

// Which pairs a BATCH covers -- a batch is what one pair warp takes with one atomic and walks pair after pair in stage 2.
// The leading pairs of a segment (its chunks that start below slot `narrow_slots`: the nearest neighbours, 8 - 30 us each in stage 2,
// and the ones that may need a hull) go out in batches of `width` pairs (32, 16, 8, 4 or 2), everything else in batches of 32:
// for a rank with few agents 32-pair batches of the nearest slots would be 0.25 - 1 ms sequential jobs of a few warps while the
// rest of the grid idles, and the inside pairs, whose hulls are the critical path of such a step, would surface late.
// Built by the host (make_batch_plan), read by the kernel from its parameters.
struct BatchPlan {   // 32-bit: lqro_create refuses ranks whose padded pair count reaches 2^31 (pair ids are ints everywhere)
  int total_a;  // padded pairs of the first segment (of everything when there is one segment)
  int total;    // padded pairs of the launch
  int width;    // pairs of a narrow batch
  int narrow_a, narrow_b;  // leading pairs of each segment in narrow batches (multiples of 32)
  int nb_a, wb_a, nb_b;    // batches: narrow ones of the first segment, wide ones of the first segment, narrow ones of the second
};
__host__ __device__ inline BatchPlan make_batch_plan(int n_local, int K, int first, int split, int narrow_slots, int width) {
  BatchPlan b;
  const int n_a = split > 0 ? split : n_local;
  b.total_a = (int)slot_major_total(n_a, K, first);
  b.total = b.total_a + (split > 0 ? (int)slot_major_total(n_local - split, K, first) : 0);
  b.width = width;
  b.narrow_a = width < 32 ? (int)slot_major_prefix(n_a, K, first, narrow_slots) : 0;
  b.narrow_b = (width < 32 && split > 0) ? (int)slot_major_prefix(n_local - split, K, first, narrow_slots) : 0;
  b.nb_a = b.narrow_a / width;
  b.wb_a = (b.total_a - b.narrow_a) / 32;
  b.nb_b = b.narrow_b / width;
  return b;
}
// batch -> first index of the enumeration and pairs of the batch; false past the end
__host__ __device__ __forceinline__ bool slot_major_batch(const BatchPlan& b, int batch, int& base, int& bw) {
  if (batch < b.nb_a) { bw = b.width; base = batch * b.width; return true; }
  int r = batch - b.nb_a;
  if (r < b.wb_a) { bw = 32; base = b.narrow_a + r * 32; return true; }
  r -= b.wb_a;
  if (r < b.nb_b) { bw = b.width; base = b.total_a + r * b.width; }
  else {
    bw = 32;
    if (r - b.nb_b > (b.total - b.total_a - b.narrow_b) / 32) return false;  // far past the end (no overflow of base)
    base = b.total_a + b.narrow_b + (r - b.nb_b) * 32;
  }
  return base < b.total;
}
