// lqro_hull.cuh -- kernels around the hull core (lqro_hull_core.cuh): the hull warps of a step (regenerate the reachable
// cloud of a queued inside pair, render it to 6 digits, hull it, nearest facet -> half-plane) and the stage kernel behind
// lqro_convex_hull / lqro_hull_escape.
#pragma once
#include "lqro_common.cuh"
#include "lqro_pair.cuh"
#include "lqro_hull_core.cuh"

__global__ void k_round6(const double* __restrict__ in, int n, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = round6_dev(in[t]);
}

// ---- kernels -------------------------------------------------------------------------------------
// shared-memory facet pool of the fast path (facets alive at once; typical LQR-obstacle hulls need < 700)
#define LQRO_HULL_SMEM_FACETS 512


// One hull slot (a main warp, in k_hull with a helper warp): its shared-memory scratch (Translate / dx / xi / step list, hot hull structures, control block)
// and its global slot (un-rounded cloud + full-size fallback pool).
struct HullWarp {
  PairWarpSmem pw;
  HullWs w;
  HullCtl* ctl;
  HullMail* mail;   // mailbox of the hull's two warps (k_hull), behind the control block
  char* hot;
  char* gslot;
};
// bytes of shared memory one hull slot needs after the block tables
__host__ __device__ inline size_t hull_warp_smem_bytes(int H, int P) {
  return hull_align16(sizeof(double) * (3 * (size_t)H + 32 + (H + 1) / 2 + 1)) + hull_hot_bytes(H * P, LQRO_HULL_SMEM_FACETS) +
         hull_align16(sizeof(HullCtl)) + hull_align16(sizeof(HullMail));
}
// the stand-alone hull kernel keeps only the ellipsoid samples (SoA, 3P doubles) in shared memory: the horizon tables are read
// from global memory (13-19 KB shared by every hull of the GPU, L1 / L2 resident; they are only needed while the cloud is
// regenerated), which is what lets a second (H = 30) or third (H = 20) hull block fit on an SM
__host__ __device__ inline size_t hull_block_smem_bytes(int H, int P, int warps = 1) {
  return hull_align16(sizeof(double) * 3 * (size_t)P) + (size_t)warps * hull_warp_smem_bytes(H, P);
}
__device__ __forceinline__ void hull_warp_setup(HullWarp& hw, char* sm, char* gslot, int H, int P) {
  double* wb = reinterpret_cast<double*>(sm);
  hw.pw.tr = wb; hw.pw.dx = wb + 3 * H; hw.pw.xi = hw.pw.dx + 16; hw.pw.klist = reinterpret_cast<int*>(hw.pw.xi + 16);
  hw.hot = sm + hull_align16(sizeof(double) * (3 * (size_t)H + 32 + (H + 1) / 2 + 1));
  hw.ctl = reinterpret_cast<HullCtl*>(hw.hot + hull_hot_bytes(H * P, LQRO_HULL_SMEM_FACETS));
  hw.mail = reinterpret_cast<HullMail*>(reinterpret_cast<char*>(hw.ctl) + hull_align16(sizeof(HullCtl)));
  hw.gslot = gslot;
  hull_carve_coords(gslot, H * P, hw.w);
  hull_carve_cold(gslot + hull_slot_cold_offset(H * P), H * P, hull_full_facets(H * P), hw.w);
  hw.pw.px = hw.w.ux; hw.pw.py = hw.w.uy; hw.pw.pz = hw.w.uz;
}

// nearest facet -> half-plane and the pair's outputs (all lanes call; lane 0 writes)
__device__ __forceinline__ void hull_finish_pair(const HullWs& w, HullCtl* ctl, const PairWarpSmem& pw, const PairOutputs& out, int pair, int n,
                                                 int* counters, long long* prof) {
  const int lane = lane_id();
  if (ctl->fail) {
    if (lane == 0) { out.status[pair] = 0x100; out.nfacets[pair] = 0; atomicAdd(counters + 3, 1); }  // no plane
    __syncwarp();
    return;
  }
  double distance, normal[3];
  hull_nearest_facet(w, *ctl, w.ux, w.uy, w.uz, pw.dx[3], pw.dx[4], pw.dx[5], &distance, normal);
  if (lane == 0 && prof) {
    long long tot = 0;
    for (int q = 0; q < 12; ++q) { atomicAdd((unsigned long long*)prof + q, (unsigned long long)ctl->prof[q]); if (q < 10) tot += ctl->prof[q]; }
    if (atomicMax(prof + 12, tot) < tot) { for (int q = 0; q < 12; ++q) prof[16 + q] = ctl->prof[q]; prof[28] = n; prof[29] = ctl->nfacets_out; }  // phases of the longest hull so far
    atomicMax(prof + 13, (long long)n);
    atomicMax(prof + 14, (long long)ctl->nfacets_out);
  }
  if (lane == 0) {
    write_plane(out.planes + 6 * (size_t)pair, pw.xi + 3, distance, normal[0], normal[1], normal[2], true);
    out.dist[pair] = distance;
    out.nfacets[pair] = ctl->nfacets_out;
    out.status[pair] = 2 | (ctl->degenerate ? 0x100 : 0);
    if (ctl->degenerate) atomicAdd(counters + 3, 1);
  }
  __syncwarp();
}
__device__ __forceinline__ void hull_render_cloud(const LqroParams& p, HullWs& w, int n) {
  for (int m = lane_id(); m < n; m += 32) {
    if (p.hull_input == 0) { w.hx[m] = round6_dev(w.ux[m]); w.hy[m] = round6_dev(w.uy[m]); w.hz[m] = round6_dev(w.uz[m]); }
    else { w.hx[m] = w.ux[m]; w.hy[m] = w.uy[m]; w.hz[m] = w.uz[m]; }
  }
  __syncwarp();
}
// the shared-memory facet pool was too small: once more with the full-size pool in the pair's global slot (rare; out of line,
// scalars only, so that the fast path keeps its pointers in registers and its shared-memory address space)
__device__ __noinline__ void hull_pair_fallback(const LqroParams& p, char* gslot, HullCtl* ctl, double* dx, double* xi, const PairOutputs& out,
                                                int pair, int n, int* counters, long long* prof) {
  HullWs wg;
  hull_ws_carve(gslot, p.HP, wg);
  hull_render_cloud(p, wg, n);
  hull_build(wg, *ctl, n, false);
  if (lane_id() == 0) atomicAdd(counters + 5, 1);
  PairWarpSmem pw;
  pw.dx = dx; pw.xi = xi;
  hull_finish_pair(wg, ctl, pw, out, pair, n, counters, prof);
}

// One queued inside-case pair: regenerate the reachable cloud (same code, same bits as k_pairs), render it to
// 6 digits, hull it, pick the nearest facet, write the plane.
// COOP: the hull's helper warp is waiting at the hull's barrier (k_hull) and is left there when this returns.
template <bool COOP = false>
__device__ __forceinline__ void hull_do_pair(const LqroParams& p, const PairTables& pt, HullWarp& hw, const double* __restrict__ x,
                                             const int* __restrict__ nbr, const PairOutputs& out, int pair, int* counters,
                                             long long* prof) {
  const int lane = lane_id();
  PairWarpSmem& pw = hw.pw;
  HullWs& w = hw.w;
  HullCtl* ctl = hw.ctl;
  const int il = pair / p.K;
  const int i = p.local_begin + il;
  const int j = nbr[pair];
  if (lane < 16) {
    const double a = x[(size_t)i * 16 + lane], b = x[(size_t)j * 16 + lane];
    pw.xi[lane] = a;
    pw.dx[lane] = a - b;
  }
  __syncwarp();
  const int n = warp_reachable_cloud(p, pt, pw);
  hull_carve_hot(hw.hot, p.HP, LQRO_HULL_SMEM_FACETS, w);   // fast path: hot structures + coordinates in shared memory
  hull_render_cloud(p, w, n);
  hull_build_t<COOP>(w, *ctl, hw.mail, n, false);
  if (ctl->fail == 2) { hull_pair_fallback(p, hw.gslot, ctl, pw.dx, pw.xi, out, pair, n, counters, prof); return; }
  hull_finish_pair(w, ctl, pw, out, pair, n, counters, prof);
}

// Hull kernel: two warps (main + helper) per queued inside pair at a time, tickets from a shared counter.  Launched twice per step of a large
// swarm:
//  * POLLING (poll != 0), on its own stream and BEFORE the pair kernel, as a concurrent consumer of the hull queue on a few SMs
//    of its own: LQRO_HULL_POLL_WARPS independent hulls (two warps each) per block, whose shared memory fills an SM, so that no pair CTA can
//    share it (a hull is a latency-bound dependent chain: next to 20 pair warps it runs 2.5x slower) -- the pair kernel's grid
//    is that many SMs short.  A polling warp takes ticket h only when the published queue length exceeds h (compare-and-swap on
//    the head), so a ticket in hand always has a producer behind it that stores the pair id within microseconds: nothing here
//    can wait for a launch that does not come.  It leaves when every pair warp of the step has signed off (producers_done ==
//    producers_total; what is still queued then goes to the wide launch at full width), or when no pair CTA has come up within
//    2 ms of its own start (launches serialised by a profiler or CUDA_LAUNCH_BLOCKING: the wide launch then does all the work).
//  * WIDE (poll == 0) after the pair kernel, one hull (two warps) per block, for whatever is left in the queue.
// Every consumer puts -1 back into the queue entry it has read: that is the "not published yet" value the next step's pollers
// rely on (the queue is filled with -1 once, when the context is created).
#define LQRO_HULL_POLL_WARPS 3   // hulls (= pairs of warps) per block of the polling launch
#ifndef LQRO_HULL_COOP
#define LQRO_HULL_COOP 1         // 0: ablation -- the helper warps leave at once, the main warp does everything (one warp per hull)
#endif
__device__ __forceinline__ unsigned long long hull_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// TWO WARPS PER HULL (lqro_hull_core.cuh, HullMail): warp 2s is the main warp of hull slot s of the block -- it takes the tickets and
// runs the build --, warp 2s + 1 its helper (planes + ridge matching next to the gather, the next insertion's horizon search next
// to the partition).  They meet at named barrier 1 + s: once per hull ("a hull follows" / "no more work"), three times per insertion.
__global__ void __launch_bounds__(64 * LQRO_HULL_POLL_WARPS, 1) k_hull(LqroParams p, LqroTables tb, const double* __restrict__ x, const int* __restrict__ nbr,
                                             PairOutputs out, HullQueue hq, int* __restrict__ work_counter, int poll,
                                             const int* __restrict__ producers_done, int producers_total, const int* __restrict__ producers_started,
                                             char* __restrict__ ws_base, size_t ws_stride, int* __restrict__ counters, long long* __restrict__ prof,
                                             const int* __restrict__ agent_class /* NULL: one class */, const LqroTables* __restrict__ class_tables) {
  extern __shared__ double smem[];
  const int lane = lane_id(), warp = threadIdx.x >> 5, hs = warp >> 1, hpb = blockDim.x >> 6;
  const bool helper = (warp & 1) != 0;
  const int bar = 1 + hs;
  PairTables pt;
  {
    const int P = p.P;
    for (int e = threadIdx.x; e < P; e += blockDim.x) { smem[e] = tb.sph[3 * e]; smem[P + e] = tb.sph[3 * e + 1]; smem[2 * P + e] = tb.sph[3 * e + 2]; }
    __syncthreads();   // the only block-wide step: from here on the hulls of a block are independent
    pt.NFt = tb.NFt; pt.T = tb.T; pt.CG = tb.CG; pt.prn = tb.prn;
    pt.sphx = smem; pt.sphy = smem + P; pt.sphz = smem + 2 * P;
  }
  const int slot = blockIdx.x * hpb + hs;
  HullWarp hw;
  hull_warp_setup(hw, reinterpret_cast<char*>(smem) + hull_align16(sizeof(double) * 3 * (size_t)p.P) + (size_t)hs * hull_warp_smem_bytes(p.H, p.P),
                  ws_base + ws_stride * (size_t)slot, p.H, p.P);
  if (helper) {
    if (!LQRO_HULL_COOP) return;
    hull_carve_hot(hw.hot, p.HP, LQRO_HULL_SMEM_FACETS, hw.w);
    while (true) {
      hull_pair_bar(bar);                     // the main warp says whether a hull follows
      if (hw.mail->hull == 0) break;
      hull_helper(hw.w, *hw.ctl, *hw.mail);
    }
    return;
  }
  int nq = 0;
  bool any = true;
  if (!poll) {
    nq = *hq.count;
    if (nq > hq.capacity) nq = hq.capacity;
    if (*(volatile int*)work_counter + slot >= nq) any = false;  // nothing can be left for this hull slot
  }
  int loaded = agent_class ? -1 : 0;
  const unsigned long long t_start = poll ? hull_globaltimer() : 0ull;
  while (any) {
    int item = -1;
    if (lane == 0) {
      if (!poll) {
        item = atomicAdd(work_counter, 1);
        if (item >= nq) { atomicSub(work_counter, 1); item = -1; }
      } else {
        while (true) {
          if (*(volatile const int*)producers_done >= producers_total) break;  // pair kernel finished: the rest goes to the wide launch
          int c = *(volatile int*)hq.count;
          if (c > hq.capacity) c = hq.capacity;
          const int h = *(volatile int*)work_counter;
          if (h < c) {
            if (atomicCAS(work_counter, h, h + 1) == h) { item = h; break; }
            continue;
          }
          if (*(volatile const int*)producers_started == 0 && hull_globaltimer() - t_start > 2000000ull) break;
          __nanosleep(256);
        }
      }
    }
    item = __shfl_sync(LQRO_FULL, item, 0);
    if (item < 0) break;
    int pair = -1;
    if (lane == 0) {
      while ((pair = *(volatile int*)(hq.items + item)) < 0) __nanosleep(32);  // published right after the slot was taken (k_pairs)
      hq.items[item] = -1;
      hw.mail->hull = 1; hw.mail->bar_id = bar;
    }
    pair = __shfl_sync(LQRO_FULL, pair, 0);
    if (LQRO_HULL_COOP) hull_pair_bar(bar);   // a hull follows: the helper enters hull_helper
    if (agent_class) {  // heterogeneous dynamics: the obstacle of pair (i, j) uses agent i's class (MAIN:1384)
      const int c = agent_class[p.local_begin + pair / p.K];
      if (c != loaded) { const LqroTables tc = class_tables[c]; pt.NFt = tc.NFt; pt.T = tc.T; pt.CG = tc.CG; pt.prn = tc.prn; loaded = c; }
    }
    hull_do_pair<LQRO_HULL_COOP != 0>(p, pt, hw, x, nbr, out, pair, counters, prof);
  }
  if (lane == 0) hw.mail->hull = 0;
  if (LQRO_HULL_COOP) hull_pair_bar(bar);     // no more work: the helper leaves
}

// Longest hulls first.  When more inside pairs are queued than hull warps can run at once (the reference's dense 32 / 64-quad
// scenarios: 100 - 500 hulls of up to H*P points per step), the wide launch works through the queue in rounds and the step ends
// with whatever the last round holds.  One hull costs about its cloud size (insertions are sequential), so the queue is put in
// order of decreasing reachable-cloud size first (longest-processing-time-first list scheduling): the long hulls start at once,
// the short ones fill the gaps.  One block; in-place bitonic sort of up to LQRO_HULL_SORT_MAX entries in shared memory, key =
// (cloud size, queue position) so the order is deterministic.  Returns at once when one round suffices.  The half-planes do not
// depend on the order.
#define LQRO_HULL_SORT_MAX 4096
__global__ void __launch_bounds__(1024, 1) k_hull_order(HullQueue hq, const int* __restrict__ work_counter, const int* __restrict__ nreach, int slots) {
  __shared__ unsigned long long key[LQRO_HULL_SORT_MAX];
  int c = *hq.count;
  if (c > hq.capacity) c = hq.capacity;
  const int h = *work_counter;
  const int m = c - h;
  if (m <= slots || m > LQRO_HULL_SORT_MAX) return;
  int np2 = 1;
  while (np2 < m) np2 <<= 1;
  for (int q = threadIdx.x; q < np2; q += blockDim.x) {
    unsigned long long k = 0ull;  // padding sorts last
    if (q < m) {
      const int pair = hq.items[h + q];
      k = ((unsigned long long)(unsigned)(nreach[pair] + 1) << 44) | ((unsigned long long)(unsigned)(LQRO_HULL_SORT_MAX - q) << 32) | (unsigned)pair;
    }
    key[q] = k;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= np2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int q = threadIdx.x; q < np2; q += blockDim.x) {
        const int partner = q ^ j;
        if (partner > q) {
          const unsigned long long a = key[q], b = key[partner];
          const bool desc = (q & k2) == 0;  // descending overall
          if (desc ? a < b : a > b) { key[q] = b; key[partner] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int q = threadIdx.x; q < m; q += blockDim.x) hq.items[h + q] = (int)(unsigned)(key[q] & 0xffffffffull);
}

// Stage kernel: hull of an arbitrary cloud (AoS pts in global memory), everything in the global slot.
// mode bit 0: apply the 6-digit rendering first.  Outputs facets in qhull order (planes nf*4, verts nf*3
// as point indices); when vrel != NULL also the nearest-facet result (escape[0] = distance,
// [1..3] = normal) and best facet.
__global__ void __launch_bounds__(32, 1) k_stage_hull(const double* __restrict__ pts, int n, int mode, char* __restrict__ ws_base, int cap_pts,
                                                   double* __restrict__ planes, int* __restrict__ verts, int maxf, int* __restrict__ info /*nf, degenerate, fail, ninserted, best*/,
                                                   const double* __restrict__ vrel, double* __restrict__ escape) {
  __shared__ HullCtl ctl;
  const int lane = lane_id();
  HullWs w;
  hull_ws_carve(ws_base, cap_pts, w);
  for (int m = lane; m < n && m < cap_pts; m += 32) {
    const double a = pts[3 * m], b = pts[3 * m + 1], c = pts[3 * m + 2];
    w.ux[m] = a; w.uy[m] = b; w.uz[m] = c;
    if (mode & 1) { w.hx[m] = round6_dev(a); w.hy[m] = round6_dev(b); w.hz[m] = round6_dev(c); }
    else { w.hx[m] = a; w.hy[m] = b; w.hz[m] = c; }
  }
  __syncwarp();
  hull_build(w, ctl, n, true);
  if (lane == 0) { info[0] = ctl.fail ? 0 : ctl.nfacets_out; info[1] = ctl.degenerate; info[2] = ctl.fail; info[3] = ctl.ninserted; info[4] = -1; }
  __syncwarp();
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
    if (lane == 0) { escape[0] = distance; escape[1] = normal[0]; escape[2] = normal[1]; escape[3] = normal[2]; info[4] = pos; }
  }
}
