// lqro_step.cuh -- the pair kernel of a step, with the hull work of the inside case overlapped.
//
// Persistent CTAs, one per SM.  Warp roles:
//
//  * PAIR warps take BATCHES of 32 pair slots (one atomic per batch) and run two stages:
//      stage 1, lane = pair (32 pairs in flight, no shuffles, table reads are warp-uniform broadcasts): load
//        x_i, x_j, form dx; per horizon step evaluate Translate_k with FMAs (approximation charged to the bound)
//        and the exact pruning rule -> 64-bit mask of surviving steps.  Pairs with no surviving step finish here
//        (status NONE, coalesced stores).
//      stage 2, warp = pair, for the lanes with a non-zero mask: Translate_k of the surviving steps in the
//        reference's operation order, candidate generation + reachable filter into the warp's cloud (first
//        LQRO_PAIR_CLOUD_CAP points in shared memory, overflow in a per-warp HBM area), GJK, half-plane
//        (outside case) or a ticket in the hull queue (inside case).
//    Batches are handed out SLOT-MAJOR (all agents' slots 0..31 first): neighbour lists are sorted by distance,
//    so the inside pairs -- the ones that need a hull -- surface early in the kernel.
//
//  * one HULL warp per CTA (when its structures fit next to the pair warps) consumes the hull queue while the
//    pair warps are still producing: qhull's insertion order makes one hull a ~1.5 ms sequential job, which
//    would otherwise sit on the critical path of the step after the pair kernel.  Whatever is still queued
//    when the producers finish is drained by the wide stand-alone k_hull launched right after.
//    Queue protocol: producers take a slot with atomicAdd(count), store the pair id, fence; consumers take
//    tickets with atomicAdd(head) and poll items[ticket] (initialised to -1 every step) until it is filled or
//    every producer warp has signed off.  Producers never wait for consumers.
#pragma once
#include "lqro_common.cuh"
#include "lqro_gjk.cuh"
#include "lqro_pair.cuh"
#include "lqro_hull.cuh"
#include "lqro_enum.cuh"  // slot-major enumeration of the pairs of a launch, batches

#ifndef LQRO_PAIR_CLOUD_CAP
#define LQRO_PAIR_CLOUD_CAP 320  // measured 192 -> 4.87 ms, 256 -> 4.82, 320 -> 4.80, 384 -> 4.78 (does not leave room for 20 warps at H = 30)
#endif
#ifndef LQRO_PAIR_MAX_WARPS
#define LQRO_PAIR_MAX_WARPS 20  // x 32 threads x 96 registers (measured: 16 -> 7.05 ms, 18 -> 7.30, 20 -> 6.84, 22 -> 7.23, 24 -> 7.22)
#endif

struct PairWarpV2 {
  double *cx, *cy, *cz;  // [LQRO_PAIR_CLOUD_CAP]
  double* tr;            // [3H]
  double* dx;            // [16]
  double* xi;            // [16]
  int* klist;            // [H]
  double* gjk;           // [LQRO_GJK_WARP_DOUBLES] simplex points + dot products of the lane-parallel sub-algorithm
};
__host__ __device__ inline size_t pair2_warp_doubles(int H) {
  return (size_t)3 * LQRO_PAIR_CLOUD_CAP + 3 * H + 32 + (H + 1) / 2 + 1 + LQRO_GJK_WARP_DOUBLES;
}

struct StepSync {
  int* batch_counter;   // next batch of pairs of this launch
  int narrow_pairs;     // = plan.narrow_a and
  int batch_width;      // = plan.width, for the instantiations of k_pairs that walk ONE segment of agents
  int* started;         // pair CTAs that have come up (the polling hull launch gives up when none does, see k_hull)
  int* hull_head;       // next hull ticket
  int* done_warps;      // pair warps that have finished
  int* counters;        // misc counters (degenerate, fallback, ...)
  long long* prof;
  BatchPlan plan;       // which pairs a batch covers (lqro_enum.cuh): narrow batches for the nearest slots of a small rank
  const int* nbr_mark;  // NULL, or the number of leading rows of the neighbour table that have arrived on the device (a word the
                        // copy engine writes behind every block of rows while this launch runs, see launch_step)
};
#ifndef LQRO_NBR_MARK_TIMEOUT_NS
#define LQRO_NBR_MARK_TIMEOUT_NS 2000000000ull  // 2 s
#endif
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

#ifdef LQRO_PAIR_PROF  // per-phase cycle counts of the pair warps (lane 0), profiling builds only
#define PAIR_PROF_DECL long long pprof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gprof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ptick = clock64();
#define PAIR_PROF(q) { const long long tn_ = clock64(); pprof[q] += tn_ - ptick; ptick = tn_; }
#define PAIR_PROF_COUNT(q, v) { pprof[q] += (v); }
#define PAIR_PROF_FLUSH if (lane == 0) { for (int q_ = 0; q_ < 8; ++q_) { atomicAdd((unsigned long long*)sy.prof + 16 + q_, (unsigned long long)pprof[q_]); atomicAdd((unsigned long long*)sy.prof + 24 + q_, (unsigned long long)gprof[q_]); } }
#define PAIR_GJK_PROF_ARG , gprof
#else
#define PAIR_PROF_DECL
#define PAIR_PROF(q)
#define PAIR_PROF_COUNT(q, v)
#define PAIR_PROF_FLUSH
#define PAIR_GJK_PROF_ARG
#endif

// HT, PT > 0: horizon / sample counts known at compile time (table offsets and loop bounds fold into immediates);
// 0: taken from the parameters.
// STREAMED: the neighbour table is still arriving while the launch runs (StepSync::nbr_mark, LqroParams::seg_split, StepSync::plan);
// the other instantiations carry none of that code -- their register allocation sits on a cliff (96 registers, 20 warps): the same
// batch arithmetic read from the BatchPlan instead of narrow_pairs / batch_width costs a 16-byte stack frame and 0.75 % of the step.
template <bool FUSED_HULL, int HT, int PT, bool STREAMED = false>
__global__ void __launch_bounds__(LQRO_PAIR_MAX_WARPS * 32, 1) k_pairs(LqroParams p, LqroTables tb, const double* __restrict__ x, const int* __restrict__ nbr,
                                                  PairOutputs out, StepSync sy, HullQueue hq, double* __restrict__ ovf_base,
                                                  size_t ovf_stride, int n_pair_warps, char* __restrict__ hull_ws, size_t hull_ws_stride,
                                                  const int* __restrict__ order /* NULL, or the owned agents of one dynamics class: p.n_local of them */) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int H = HT > 0 ? HT : p.H, P = PT > 0 ? PT : p.P;
  PairTables t;
  load_block_tables(smem, tb, H, P);
  {
    t.NFt = smem; t.T = smem + 16 * 3 * H; t.sphx = t.T + 9 * H; t.sphy = t.sphx + P; t.sphz = t.sphy + P;
    t.CG = t.sphz + P; t.prn = t.CG + 9 * H;
  }
  const int total_pair_warps = n_pair_warps * gridDim.x;

  if (FUSED_HULL && warp >= n_pair_warps) {
    // ------------------------------------------------------------------ hull warp (consumer)
    HullWarp hw;
    char* sm = reinterpret_cast<char*>(smem) +
               hull_align16(sizeof(double) * (pair_block_table_doubles(H, P) + (size_t)n_pair_warps * pair2_warp_doubles(H)));
    hull_warp_setup(hw, sm, hull_ws + hull_ws_stride * (size_t)blockIdx.x, H, P);
    while (true) {
      int ticket = -1;
      if (lane == 0) {
        // stop taking tickets once every producer has signed off: the stand-alone k_hull drains the rest
        if (*(volatile int*)sy.done_warps < total_pair_warps) ticket = atomicAdd(sy.hull_head, 1);
      }
      ticket = __shfl_sync(LQRO_FULL, ticket, 0);
      if (ticket < 0 || ticket >= hq.capacity) break;
      int pair = -1;
      if (lane == 0) {
        while (true) {
          pair = *(volatile int*)(hq.items + ticket);
          if (pair >= 0) break;
          if (*(volatile int*)sy.done_warps >= total_pair_warps) {  // all items are published by now
            pair = *(volatile int*)(hq.items + ticket);
            break;
          }
          __nanosleep(400);
        }
      }
      pair = __shfl_sync(LQRO_FULL, pair, 0);
      if (pair < 0) {  // ticket beyond the final tail: hand it back so k_hull's scan stays contiguous
        break;
      }
      hull_do_pair(p, t, hw, x, nbr, out, pair, sy.counters, sy.prof);
      if (lane == 0) hq.items[ticket] = -2 - pair;  // done marker (k_hull skips negative entries)
    }
    return;
  }

  // -------------------------------------------------------------------- pair warps (producers)
  PairWarpV2 w;
  {
    double* wb = smem + pair_block_table_doubles(H, P) + (size_t)warp * pair2_warp_doubles(H);
    w.cx = wb; w.cy = wb + LQRO_PAIR_CLOUD_CAP; w.cz = wb + 2 * LQRO_PAIR_CLOUD_CAP;
    w.tr = wb + 3 * LQRO_PAIR_CLOUD_CAP; w.dx = w.tr + 3 * H; w.xi = w.dx + 16;
    w.klist = reinterpret_cast<int*>(w.xi + 16);
    w.gjk = w.xi + 16 + (H + 1) / 2 + 1;
  }
  CloudView cl;
  cl.sx = w.cx; cl.sy = w.cy; cl.sz = w.cz; cl.cap = LQRO_PAIR_CLOUD_CAP;
  {
    double* ov = ovf_base + ovf_stride * (size_t)(blockIdx.x * n_pair_warps + warp);
    const size_t no = (size_t)((H * P) > LQRO_PAIR_CLOUD_CAP ? (H * P) - LQRO_PAIR_CLOUD_CAP : 0);
    cl.gx = ov; cl.gy = ov + no; cl.gz = ov + 2 * no;
  }
  const long long total64 = slot_major_total(p.n_local, p.K, p.first_slots);
  const int nbw = sy.batch_width;
  const int narrow_batches = sy.narrow_pairs / nbw;
  bool nbr_complete = !STREAMED || sy.nbr_mark == nullptr;  // else: rows [0, *nbr_mark) of the neighbour table have arrived
  if (threadIdx.x == 0 && sy.started) atomicAdd(sy.started, 1);
  const int n3 = 3 * H;
  const unsigned long long allsteps = (H >= 64) ? ~0ull : ((1ull << H) - 1ull);
  PAIR_PROF_DECL
  while (true) {
    int bw = 0, base = -1;
    if (STREAMED) {  // two segments of agents (BatchPlan)
      if (lane == 0 && !slot_major_batch(sy.plan, atomicAdd(sy.batch_counter, 1), base, bw)) base = -1;
      base = __shfl_sync(LQRO_FULL, base, 0);
      bw = __shfl_sync(LQRO_FULL, bw, 0);
      if (base < 0) break;
    } else {  // one segment: narrow batches, then wide ones
      int batch = 0;
      if (lane == 0) batch = atomicAdd(sy.batch_counter, 1);
      batch = __shfl_sync(LQRO_FULL, batch, 0);
      bw = batch < narrow_batches ? nbw : 32;
      const long long base64 = batch < narrow_batches ? (long long)batch * nbw : (long long)sy.narrow_pairs + (long long)(batch - narrow_batches) * 32;
      if (base64 >= total64) break;
      base = (int)base64;
    }
    // ---------------- stage 1: lane = pair
    int il = 0, kk = 0;
    const bool have = lane < bw && (STREAMED ? slot_major_pair(p, sy.plan.total_a, base + lane, il, kk)
                                          : slot_major_pair_n<long long>(p.n_local, p.K, p.first_slots, (long long)base + lane, il, kk));
    if (STREAMED && !nbr_complete) {
      // the neighbour table is being uploaded next to this launch (launch_step: rows in order, the mark after each block, both
      // written by the copy engine): wait for the rows of this batch.  A mark that never comes (a failed copy) ends the wait
      // after LQRO_NBR_MARK_TIMEOUT_NS and raises counter 9, which the host turns into an error.
      const int need = __reduce_max_sync(LQRO_FULL, have ? il + 1 : 0);
      int got = ld_acquire_sys(sy.nbr_mark);
      if (got < need) {
        const unsigned long long t0 = global_timer_ns();
        while ((got = ld_acquire_sys(sy.nbr_mark)) < need) {
          // given up once (by this warp or any other): the step is lost, nobody waits again -- the launch ends within one timeout
          if (*(volatile int*)(sy.counters + 9) != 0 || global_timer_ns() - t0 > LQRO_NBR_MARK_TIMEOUT_NS) {
            if (lane == 0) atomicExch(sy.counters + 9, 1);
            got = p.n_local;
            break;
          }
          __nanosleep(200);
        }
      }
      nbr_complete = got >= p.n_local;
    }
    if (have && order) il = order[il];
    const int pair = have ? il * p.K + kk : -1;
    int i = 0, j = -1;
    unsigned long long kmask = 0ull;
    if (have) {
      i = p.local_begin + il;
      j = nbr[pair];
      if (j >= 0 && j != i && j < p.N) {  // the reference skips i == j (MAIN:1379)
        if (!p.prune) {
          kmask = allsteps;
        } else {
          double dx[16];
          const double2* xi2 = reinterpret_cast<const double2*>(x + (size_t)i * 16);
          const double2* xj2 = reinterpret_cast<const double2*>(x + (size_t)j * 16);
          double dx1 = 0.0;
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const double2 a = xi2[m], b = xj2[m];
            dx[2 * m] = a.x - b.x; dx[2 * m + 1] = a.y - b.y;
            dx1 += fabs(dx[2 * m]) + fabs(dx[2 * m + 1]);
          }
          const double vr0 = dx[3], vr1 = dx[4], vr2 = dx[5];
          // state differences beyond position and velocity that are exactly zero (agents in the same attitude / rotor state, as
          // every pair of the synthetic swarms) add exactly nothing to the sums: skip their terms
          bool tail0 = true;
#pragma unroll
          for (int m = 6; m < 16; ++m) tail0 = tail0 && dx[m] == 0.0;
          for (int k = 0; k < H; ++k) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            const double* nf = t.NFt + 3 * k;
#pragma unroll
            for (int m = 0; m < 6; ++m) {
              t0 = fma(nf[m * n3], dx[m], t0);
              t1 = fma(nf[m * n3 + 1], dx[m], t1);
              t2 = fma(nf[m * n3 + 2], dx[m], t2);
            }
            if (!tail0) {
#pragma unroll
              for (int m = 6; m < 16; ++m) {
                t0 = fma(nf[m * n3], dx[m], t0);
                t1 = fma(nf[m * n3 + 1], dx[m], t1);
                t2 = fma(nf[m * n3 + 2], dx[m], t2);
              }
            }
            // |fused sum - reference's rounded sum| <= 2 * 16 * 2^-53 * sum |NF||dx| <= 1e-14 * max|NF_k| * |dx|_1
            const double tr_err = 1.0e-14 * t.prn[LQRO_PRN * k + 7] * dx1;
            if (!step_prunable_v(p, t, k, t0, t1, t2, tr_err, vr0, vr1, vr2)) kmask |= 1ull << k;
          }
        }
      }
      if (kmask == 0ull) { out.status[pair] = 0; out.nreach[pair] = 0; }
    }
    // ---------------- stage 2: warp = pair
    unsigned act = __ballot_sync(LQRO_FULL, kmask != 0ull);
    PAIR_PROF(0)
    PAIR_PROF_COUNT(6, __popc(act))
    // the two states of a pair come from L2 (~1 us): those of the NEXT stage-2 pair are requested while this one is worked on
    double pfa = 0.0, pfb = 0.0;
    if (act) {
      const int s0 = __ffs(act) - 1;
      const int i0 = __shfl_sync(LQRO_FULL, i, s0), j0 = __shfl_sync(LQRO_FULL, j, s0);
      if (lane < 16) { pfa = x[(size_t)i0 * 16 + lane]; pfb = x[(size_t)j0 * 16 + lane]; }
    }
    while (act) {
      const int src = __ffs(act) - 1;
      act &= act - 1;
      const int pr = __shfl_sync(LQRO_FULL, pair, src);
      const unsigned long long km = __shfl_sync(LQRO_FULL, kmask, src);
      if (lane < 16) {
        w.xi[lane] = pfa;
        w.dx[lane] = pfa - pfb;  // xInit1 - xInit2 (MAIN:685); entries 3..5 are also v_rel (MAIN:698-700)
      }
      if (act) {
        const int s1 = __ffs(act) - 1;
        const int i1 = __shfl_sync(LQRO_FULL, i, s1), j1 = __shfl_sync(LQRO_FULL, j, s1);
        if (lane < 16) { pfa = x[(size_t)i1 * 16 + lane]; pfb = x[(size_t)j1 * 16 + lane]; }
      }
      const int nk = __popcll(km);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        if (k < H && ((km >> k) & 1ull)) w.klist[__popcll(km & ((1ull << k) - 1ull))] = k;
      }
      __syncwarp();
      // Translate_k of the surviving steps in the reference's operation order, then candidates + reachable filter into the
      // warp's cloud (lqro_pair.cuh)
      pair_translate_steps(t.NFt, w.dx, w.klist, nk, n3, w.tr);
      __syncwarp();
      PAIR_PROF(1)
      const double vr0 = w.dx[3], vr1 = w.dx[4], vr2 = w.dx[5];
      const int count = pair_scan_cloud<PT>(t, w.tr, w.klist, nk, P, vr0, vr1, vr2, p.r2, p.inv_r2, p.r2lo, p.r2hi, cl);
      PAIR_PROF(2)
      const int n = count;
      PAIR_PROF_COUNT(7, n)
      if (n <= 4) {  // "GJK errors with less than 4", MAIN:1393
        if (lane == 0) { out.status[pr] = 0; out.nreach[pr] = n; }
        __syncwarp();
        continue;
      }
      double gw1[3], gw2[3];
#if LQRO_GJK_LANES
      const double sq = gjk_warp_lanes(vr0, vr1, vr2, cl, n, w.gjk, gw1, gw2, nullptr PAIR_GJK_PROF_ARG);
#else
      const double sq = gjk_warp(vr0, vr1, vr2, cl, n, gw1, gw2, nullptr);
#endif
      PAIR_PROF(3)
      {
        const double distance = sqrt(sq);                             // MAIN:741
        const bool inside = distance < 0.0001 && distance > -0.0001;  // MAIN:752-757
        if (!inside) {
          // lanes 0..2 each finish one component: normal (MAIN:743-745), then createHalfPlanes (MAIN:1085-1098)
          // after `distance *= 0.5` (MAIN:1401), narrowed to float as Vector3(double,double,double) does
          if (lane < 3) {
            const double a = lane == 0 ? gw1[0] : lane == 1 ? gw1[1] : gw1[2];
            const double b = lane == 0 ? gw2[0] : lane == 1 ? gw2[1] : gw2[2];
            const double nc = (a - b) / distance;
            const double md = -1.0 * (distance * 0.5);
            float* pl = out.planes + 6 * (size_t)pr;
            pl[lane] = (float)(w.xi[3 + lane] + md * nc);
            pl[3 + lane] = (float)nc;
          }
          if (lane == 3) { out.nreach[pr] = n; out.dist[pr] = distance; out.status[pr] = 1; }
        } else if (lane == 0) {
          out.nreach[pr] = n;
          const int slot = atomicAdd(hq.count, 1);
          out.status[pr] = slot < hq.capacity ? 2 : 0x200;  // 2: plane filled in by a hull warp / k_hull; 0x200: queue full, no plane
          __threadfence();
          if (slot < hq.capacity) { *(volatile int*)(hq.items + slot) = pr; }
        }
      }
      __syncwarp();
      PAIR_PROF(4)
    }
  }
  PAIR_PROF_FLUSH
  if (lane == 0) { __threadfence(); atomicAdd(sy.done_warps, 1); }
}
