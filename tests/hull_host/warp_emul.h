// tests/hull_host/warp_emul.h -- TEST INFRASTRUCTURE ONLY.
// A 32-lane lockstep emulation of the CUDA warp-level primitives the hull kernel uses, so that the KERNEL SOURCE
// (lqr_obstacles_b200/csrc/lqro_hull_core.cuh) can be compiled as plain C++ and run on a machine without a GPU:
// every lane is a real host thread, every warp collective (__shfl_sync, __ballot_sync, __match_any_sync,
// __reduce_*_sync, __syncwarp) is a rendezvous of the 32 threads.  A collective that not every lane reaches
// deadlocks here exactly as a full-mask collective under divergence is undefined on the device, so the emulation
// also checks the convergence discipline of the kernel.  The product (liblqro.so) never contains this.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __align__(n) alignas(n)
#define LQRO_FULL 0xffffffffu
#define LQRO_HULL_HOST_EMUL 1

struct WarpEmu {
  std::barrier<> bar{32};
  uint64_t slot[2][32];
  WarpEmu() { memset(slot, 0, sizeof slot); }
};
static thread_local int tl_lane = 0;
static thread_local unsigned tl_phase = 0;
static thread_local WarpEmu* tl_warp = nullptr;
// block of several emulated warps: one barrier all of their threads meet (stands in for a named barrier, bar.sync id, 32 * warps)
static thread_local std::barrier<>* tl_block_bar = nullptr;
static inline void emu_block_barrier() { tl_block_bar->arrive_and_wait(); }

static inline int lane_id() { return tl_lane; }
static inline long long clock64() { return 0; }

// one rendezvous: publish v, wait, return the bank every lane published into (double buffered: a lane can be at most
// one collective ahead of the slowest one, and it only overwrites the bank of the collective before the previous one)
static inline const uint64_t* emu_exchange(uint64_t v) {
  WarpEmu* w = tl_warp;
  const unsigned b = tl_phase++ & 1u;
  w->slot[b][tl_lane] = v;
  w->bar.arrive_and_wait();
  return w->slot[b];
}
static inline void __syncwarp(unsigned = LQRO_FULL) { tl_phase++; tl_warp->bar.arrive_and_wait(); }
static inline void __threadfence() {}

template <typename T> static inline uint64_t emu_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T emu_val(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { const uint64_t* s = emu_exchange(emu_bits(v)); return emu_val<T>(s[src & 31]); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { const uint64_t* s = emu_exchange(emu_bits(v)); return emu_val<T>(s[(tl_lane ^ m) & 31]); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { const uint64_t* s = emu_exchange(emu_bits(v)); return emu_val<T>(s[tl_lane >= d ? tl_lane - d : tl_lane]); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { const uint64_t* s = emu_exchange(emu_bits(v)); return emu_val<T>(s[tl_lane + d < 32 ? tl_lane + d : tl_lane]); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  const uint64_t* s = emu_exchange(pred ? 1u : 0u);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] & 1u) << i;
  return m;
}
static inline int __any_sync(unsigned, int pred) { return __ballot_sync(LQRO_FULL, pred) != 0u; }
static inline int __all_sync(unsigned, int pred) { return __ballot_sync(LQRO_FULL, pred) == 0xffffffffu; }
template <typename T> static inline unsigned __match_any_sync(unsigned, T v) {
  const uint64_t mine = emu_bits(v);
  const uint64_t* s = emu_exchange(mine);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] == mine) << i;
  return m;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { const uint64_t* s = emu_exchange(v); unsigned r = 0; for (int i = 0; i < 32; ++i) if ((unsigned)s[i] > r) r = (unsigned)s[i]; return r; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { const uint64_t* s = emu_exchange(v); unsigned r = 0xffffffffu; for (int i = 0; i < 32; ++i) if ((unsigned)s[i] < r) r = (unsigned)s[i]; return r; }
static inline int __reduce_max_sync(unsigned, int v) { const uint64_t* s = emu_exchange(emu_bits(v)); int r = emu_val<int>(s[0]); for (int i = 1; i < 32; ++i) { const int q = emu_val<int>(s[i]); if (q > r) r = q; } return r; }
static inline int __reduce_min_sync(unsigned, int v) { const uint64_t* s = emu_exchange(emu_bits(v)); int r = emu_val<int>(s[0]); for (int i = 1; i < 32; ++i) { const int q = emu_val<int>(s[i]); if (q < r) r = q; } return r; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { const uint64_t* s = emu_exchange(v); unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (unsigned)s[i]; return r; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { const uint64_t* s = emu_exchange(v); unsigned r = 0; for (int i = 0; i < 32; ++i) r += (unsigned)s[i]; return r; }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
// n-th (1-based) set bit of mask at or above bit `base` (offset > 0), as the device intrinsic
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  for (unsigned b = base; b < 32; ++b) if ((mask >> b) & 1u) { if (--offset == 0) return b; }
  return 0xffffffffu;
}

template <typename T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicMin(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }

// run fn(warp, lane) on nwarps x 32 host threads; the lanes of a warp rendezvous in the warp collectives, all threads in emu_block_barrier()
template <typename Fn> static inline void emu_run_block(int nwarps, Fn fn) {
  std::vector<WarpEmu> warps(nwarps);
  std::barrier<> bar(32 * nwarps);
  std::vector<std::thread> th;
  for (int wi = 0; wi < nwarps; ++wi)
    for (int l = 0; l < 32; ++l)
      th.emplace_back([&warps, &bar, wi, l, &fn]() { tl_lane = l; tl_phase = 0; tl_warp = &warps[wi]; tl_block_bar = &bar; fn(wi, l); });
  for (auto& t : th) t.join();
}
// run fn(lane) on 32 lockstep host threads
template <typename Fn> static inline void emu_run_warp(Fn fn) {
  WarpEmu warp;
  std::vector<std::thread> th;
  for (int l = 0; l < 32; ++l)
    th.emplace_back([&warp, l, &fn]() { tl_lane = l; tl_phase = 0; tl_warp = &warp; fn(l); });
  for (auto& t : th) t.join();
}
