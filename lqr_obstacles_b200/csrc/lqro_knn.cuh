// lqro_knn.cuh -- neighbour search on the device.
//
// What it computes: for every owned agent i the K other agents with the smallest squared distance
//   distSq = (dx*dx + dy*dy) + dz*dz,  d = p_j - p_i           (absSq of the offset, AGENT:160)
// ordered by (distSq, j) ascending -- the order RVO's Agent::insertAgentNeighbor (AGENT:153-174) builds by
// insertion, with equal distances resolved to the lower index; -1 pads the list when fewer than K others exist.
// The result is a pure function of the positions (the candidate visiting order cannot change it), so it is
// compared bit for bit with the brute-force oracle (oracle/lqro_oracle.cpp: lqro_o_knn).
//
// How: uniform grid over the bounding box of all N positions.
//   k_knn_bbox     block min/max -> atomics on order-preserving 64-bit keys
//   k_knn_setup    one thread: cell size ~ the expected K-th neighbour distance, grid dimensions
//   k_knn_count    cell of every agent, per-cell counts
//   k_knn_scan     one block: exclusive scan of the counts (<= LQRO_KNN_MAX_CELLS)
//   k_knn_scatter  agent ids and positions in cell order (x fastest => one row of cells is one contiguous range)
//   k_knn_query    one warp per owned agent: shells of cells of growing Chebyshev radius r; candidates closer
//                  than the current K-th best are staged in shared memory and merged into a register-resident
//                  sorted top-K (bitonic sort + bitonic merge, 32*E entries, E per lane); the search ends when
//                  the K-th best is strictly inside the distance to the nearest unsearched cell face.
#pragma once
#include "lqro_common.cuh"

#define LQRO_KNN_MAX_CELLS (1 << 18)
#define LQRO_KNN_STAGE 32  // extra staging entries beyond one merge batch
#define LQRO_KNN_FAST_CAP 128  // staging entries of the histogram-seeded fast path
// shared memory per warp: staging keys (double + int) and the 32-bin histogram
__host__ __device__ inline size_t knn_warp_smem_bytes(int E) {
  const int cap = 32 * E + LQRO_KNN_STAGE > LQRO_KNN_FAST_CAP ? 32 * E + LQRO_KNN_STAGE : LQRO_KNN_FAST_CAP;
  return (size_t)cap * (sizeof(double) + sizeof(int)) + 32 * sizeof(int);
}

struct KnnGrid {
  double org[3];
  double h, inv_h;
  int dim[3];
  int ncell;
};

__device__ __forceinline__ unsigned long long knn_key(double v) {
  const long long b = __double_as_longlong(v + 0.0);
  return (unsigned long long)b ^ ((unsigned long long)(b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double knn_unkey(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k ^ 0x8000000000000000ull) : ~k));
}

// bb[0..2] = min keys (initialised to ~0), bb[3..5] = max keys (initialised to 0)
__global__ void __launch_bounds__(256) k_knn_bbox(const double* __restrict__ x, int N, unsigned long long* __restrict__ bb) {
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const unsigned long long k = knn_key(x[(size_t)i * LQRO_XD + d]);
      lo[d] = k < lo[d] ? k : lo[d];
      hi[d] = k > hi[d] ? k : hi[d];
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long a = __shfl_xor_sync(LQRO_FULL, lo[d], off), b = __shfl_xor_sync(LQRO_FULL, hi[d], off);
      lo[d] = a < lo[d] ? a : lo[d];
      hi[d] = b > hi[d] ? b : hi[d];
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(bb + d, lo[d]); atomicMax(bb + 3 + d, hi[d]); }
  }
}

__global__ void k_knn_setup(const unsigned long long* __restrict__ bb, int N, int K, KnnGrid* __restrict__ g) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double lo[3], ext[3], vol = 1.0, emax = 0.0;
  int nd = 0;
  for (int d = 0; d < 3; ++d) {
    lo[d] = knn_unkey(bb[d]);
    ext[d] = knn_unkey(bb[3 + d]) - lo[d];
    if (ext[d] > emax) emax = ext[d];
  }
  for (int d = 0; d < 3; ++d) if (ext[d] > 1e-9 * emax) { vol *= ext[d]; ++nd; }
  // cell size: a bit more than the radius that holds K neighbours at the mean density (so that, away from the
  // swarm's boundary, the first shell already decides); flat / linear swarms use their own dimension
  double h = emax > 0 ? emax : 1.0;
  if (nd > 0 && N > 1) {
    const double per = vol * (double)(K + 1) / (double)N;  // volume holding K+1 agents
    const double rk = nd == 3 ? cbrt(per * 0.2387324146378430) : nd == 2 ? sqrt(per * 0.3183098861837907) : per * 0.5;
    h = 1.1 * rk;
  }
  if (!(h > 0)) h = 1.0;
  int dim[3];
  while (true) {
    long long nc = 1;
    for (int d = 0; d < 3; ++d) { dim[d] = (int)(ext[d] / h) + 1; nc *= dim[d]; }
    if (nc <= LQRO_KNN_MAX_CELLS) break;
    h *= 1.25;
  }
  for (int d = 0; d < 3; ++d) { g->org[d] = lo[d]; g->dim[d] = dim[d]; }
  g->h = h;
  g->inv_h = 1.0 / h;
  g->ncell = dim[0] * dim[1] * dim[2];
}

__device__ __forceinline__ void knn_cell_of(const KnnGrid& g, double px, double py, double pz, int& cx, int& cy, int& cz) {
  cx = (int)((px - g.org[0]) * g.inv_h); cy = (int)((py - g.org[1]) * g.inv_h); cz = (int)((pz - g.org[2]) * g.inv_h);
  cx = cx < 0 ? 0 : (cx >= g.dim[0] ? g.dim[0] - 1 : cx);
  cy = cy < 0 ? 0 : (cy >= g.dim[1] ? g.dim[1] - 1 : cy);
  cz = cz < 0 ? 0 : (cz >= g.dim[2] ? g.dim[2] - 1 : cz);
}

__global__ void __launch_bounds__(256) k_knn_count(const double* __restrict__ x, int N, const KnnGrid* __restrict__ gp, int* __restrict__ cell,
                                                   int* __restrict__ count) {
  const KnnGrid g = *gp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int cx, cy, cz;
    knn_cell_of(g, x[(size_t)i * LQRO_XD], x[(size_t)i * LQRO_XD + 1], x[(size_t)i * LQRO_XD + 2], cx, cy, cz);
    const int c = (cz * g.dim[1] + cy) * g.dim[0] + cx;
    cell[i] = c;
    atomicAdd(count + c, 1);
  }
}

// exclusive scan of count[0..ncell) -> start[0..ncell], one block of 1024 threads; also clears count (reused as fill)
__global__ void __launch_bounds__(1024) k_knn_scan(const KnnGrid* __restrict__ gp, int* __restrict__ count, int* __restrict__ start) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int n = gp->ncell;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? count[i] : 0;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(LQRO_FULL, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int s = wsum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(LQRO_FULL, s, off);
        if (lane >= off) s += t;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int prefix = carry + (warp > 0 ? wsum[warp - 1] : 0) + incl - v;
    if (i < n) { start[i] = prefix; count[i] = 0; }
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n] = carry;
}

__global__ void __launch_bounds__(256) k_knn_scatter(const double* __restrict__ x, int N, const int* __restrict__ cell, const int* __restrict__ start,
                                                     int* __restrict__ fill, int* __restrict__ sid, double* __restrict__ spos) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int c = cell[i];
    const int pos = start[c] + atomicAdd(fill + c, 1);
    sid[pos] = i;
    spos[(size_t)pos * 3] = x[(size_t)i * LQRO_XD];
    spos[(size_t)pos * 3 + 1] = x[(size_t)i * LQRO_XD + 1];
    spos[(size_t)pos * 3 + 2] = x[(size_t)i * LQRO_XD + 2];
  }
}

// ---- sorted top-(32*E) of (distSq, index) keys, E entries per lane; entry e = j*32 + lane ------------------------
__device__ __forceinline__ bool knn_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

template <int E>
struct KnnTop {
  double d[E];
  int i[E];
};

// compare-exchange of entry (j, lane) with its partner at entry distance `dist` inside a bitonic network;
// `asc` = this entry's block sorts ascending
template <int E>
__device__ __forceinline__ void knn_cx(KnnTop<E>& t, int lane, int dist, int kblock) {
  if (dist < 32) {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const double od = __shfl_xor_sync(LQRO_FULL, t.d[j], dist);
      const int oi = __shfl_xor_sync(LQRO_FULL, t.i[j], dist);
      const int idx = j * 32 + lane;
      const bool asc = (idx & kblock) == 0;
      const bool lower = (lane & dist) == 0;          // this entry has the smaller index of the two
      const bool other_less = knn_less(od, oi, t.d[j], t.i[j]);
      // keep the minimum when (lower == asc), the maximum otherwise
      const bool take = (lower == asc) ? other_less : !other_less && !(od == t.d[j] && oi == t.i[j]);
      if (take) { t.d[j] = od; t.i[j] = oi; }
    }
  } else {
    const int jd = dist >> 5;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      if ((j & jd) == 0 && (j | jd) < E) {
        const int j2 = j | jd;
        const int idx = j * 32 + lane;
        const bool asc = (idx & kblock) == 0;
        const bool second_less = knn_less(t.d[j2], t.i[j2], t.d[j], t.i[j]);
        if (second_less == asc) {
          const double td = t.d[j]; t.d[j] = t.d[j2]; t.d[j2] = td;
          const int ti = t.i[j]; t.i[j] = t.i[j2]; t.i[j2] = ti;
        }
      }
    }
  }
}

template <int E>
__device__ __forceinline__ void knn_sort(KnnTop<E>& t, int lane) {  // ascending bitonic sort of the 32*E entries
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1)
#pragma unroll
    for (int dist = k >> 1; dist > 0; dist >>= 1) knn_cx<E>(t, lane, dist, k == 32 * E ? 0x40000000 : k);
}

// top <- the 32*E smallest of (top U batch); both sorted ascending on entry, top sorted ascending on exit
template <int E>
__device__ __forceinline__ void knn_merge(KnnTop<E>& top, const KnnTop<E>& batch, int lane) {
#pragma unroll
  for (int j = 0; j < E; ++j) {
    // entry M-1-idx of the batch lives in register E-1-j of lane 31-lane
    const double od = __shfl_sync(LQRO_FULL, batch.d[E - 1 - j], 31 - lane);
    const int oi = __shfl_sync(LQRO_FULL, batch.i[E - 1 - j], 31 - lane);
    if (knn_less(od, oi, top.d[j], top.i[j])) { top.d[j] = od; top.i[j] = oi; }
  }
#pragma unroll
  for (int dist = 16 * E; dist > 0; dist >>= 1) knn_cx<E>(top, lane, dist, 0x40000000);  // bitonic -> ascending
}

struct KnnStage {  // per-warp staging area in shared memory
  double* d;  // [32*E + LQRO_KNN_STAGE]
  int* i;
};

template <int E>
__device__ __forceinline__ void knn_flush(KnnTop<E>& top, KnnStage& st, int& nst, int lane) {
  if (nst == 0) return;
  KnnTop<E> b;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const int e = j * 32 + lane;
    if (e < nst) { b.d[j] = st.d[e]; b.i[j] = st.i[e]; }
    else { b.d[j] = __longlong_as_double(0x7ff0000000000000ll); b.i[j] = 0x7fffffff; }
  }
  __syncwarp();
  // entries beyond one batch move to the front
  const int rest = nst > 32 * E ? nst - 32 * E : 0;
  double rd = 0; int ri = 0;
  if (lane < rest) { rd = st.d[32 * E + lane]; ri = st.i[32 * E + lane]; }
  __syncwarp();
  if (lane < rest) { st.d[lane] = rd; st.i[lane] = ri; }
  nst = rest;
  knn_sort<E>(b, lane);
  knn_merge<E>(top, b, lane);
  __syncwarp();
}

// One warp per owned agent.  nbr_out [n_local][K].
template <int E>
__global__ void __launch_bounds__(128) k_knn_query(const double* __restrict__ x, int local_begin, int n_local, int K, const KnnGrid* __restrict__ gp,
                                                   const int* __restrict__ start, const int* __restrict__ sid, const double* __restrict__ spos,
                                                   int* __restrict__ nbr_out) {
  extern __shared__ double knn_smem[];
  const KnnGrid g = *gp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  constexpr int CAP = 32 * E + LQRO_KNN_STAGE > LQRO_KNN_FAST_CAP ? 32 * E + LQRO_KNN_STAGE : LQRO_KNN_FAST_CAP;
  KnnStage st;
  st.d = knn_smem + (size_t)warp * CAP;
  st.i = reinterpret_cast<int*>(knn_smem + (size_t)wpb * CAP) + (size_t)warp * CAP;
  int* hist = reinterpret_cast<int*>(knn_smem + (size_t)wpb * CAP) + (size_t)wpb * CAP + (size_t)warp * 32;
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  for (int il = blockIdx.x * wpb + warp; il < n_local; il += gridDim.x * wpb) {
    const int i = local_begin + il;
    const double px = x[(size_t)i * LQRO_XD], py = x[(size_t)i * LQRO_XD + 1], pz = x[(size_t)i * LQRO_XD + 2];
    int cx, cy, cz;
    knn_cell_of(g, px, py, pz, cx, cy, cz);
    // ---- fast path: exact threshold from a histogram of the 3x3x3 block, one collection pass, rank selection.
    // Pass 1 bins the squared distances of the block's agents (32 bins up to (1.5 h)^2) and takes the first bin
    // edge T below which at least K other agents lie; if T is inside the distance to the nearest unsearched face,
    // the K nearest are exactly the K smallest of {dsq < T} -- typically K + a few -- which pass 2 stages in shared
    // memory; every staged key then counts the keys below it and lands at that position (sorted output, no merge).
    // Anything unusual (too few agents nearby, too many below T, block not covering T) falls through to the general
    // shell search below.
    {
      const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.dim[2] - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, g.dim[1] - 1);
      const int c0 = max(cx - 1, 0), c1 = min(cx + 1, g.dim[0] - 1);
      const double rm2 = 2.25 * g.h * g.h, inv_w = 32.0 / rm2;
      hist[lane] = 0;
      __syncwarp();
      for (int zz = z0; zz <= z1; ++zz)
        for (int yy = y0; yy <= y1; ++yy) {
          const int rowbase = (zz * g.dim[1] + yy) * g.dim[0];
          const int s0 = start[rowbase + c0], s1 = start[rowbase + c1 + 1];
          for (int q = s0 + lane; q < s1; q += 32) {
            const double dx = spos[(size_t)q * 3] - px, dy = spos[(size_t)q * 3 + 1] - py, dz = spos[(size_t)q * 3 + 2] - pz;
            const double dsq = dx * dx + dy * dy + dz * dz;
            if (sid[q] != i) { const double b = dsq * inv_w; atomicAdd(&hist[b < 31.0 ? (int)b : 31], 1); }
          }
        }
      __syncwarp();
      int cum = hist[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(LQRO_FULL, cum, off);
        if (lane >= off) cum += t;
      }
      const unsigned reach = __ballot_sync(LQRO_FULL, cum >= K) & 0x7fffffffu;  // bin 31 is the catch-all
      bool done = false;
      if (reach) {
        const int B = __ffs(reach) - 1;
        // dsq * inv_w < B + 1 up to one rounding  =>  dsq < T
        const double T = (double)(B + 1) * (rm2 / 32.0) * (1.0 + 1.0e-12);
        double cov = INF;  // distance to the nearest face of the block with cells behind it
        if (cx - 1 > 0) cov = fmin(cov, px - (g.org[0] + (double)(cx - 1) * g.h));
        if (cx + 1 < g.dim[0] - 1) cov = fmin(cov, (g.org[0] + (double)(cx + 2) * g.h) - px);
        if (cy - 1 > 0) cov = fmin(cov, py - (g.org[1] + (double)(cy - 1) * g.h));
        if (cy + 1 < g.dim[1] - 1) cov = fmin(cov, (g.org[1] + (double)(cy + 2) * g.h) - py);
        if (cz - 1 > 0) cov = fmin(cov, pz - (g.org[2] + (double)(cz - 1) * g.h));
        if (cz + 1 < g.dim[2] - 1) cov = fmin(cov, (g.org[2] + (double)(cz + 2) * g.h) - pz);
        const double safe = cov * (1.0 - 1.0e-9) - 1.0e-9 * g.h;
        if (cov == INF || (safe > 0 && T < safe * safe)) {
          int c = 0;
          bool fits = true;
          for (int zz = z0; zz <= z1 && fits; ++zz)
            for (int yy = y0; yy <= y1 && fits; ++yy) {
              const int rowbase = (zz * g.dim[1] + yy) * g.dim[0];
              const int s0 = start[rowbase + c0], s1 = start[rowbase + c1 + 1];
              for (int b0 = s0; b0 < s1; b0 += 32) {
                const int q = b0 + lane;
                bool keep = false;
                double dsq = 0;
                int j = 0;
                if (q < s1) {
                  j = sid[q];
                  const double dx = spos[(size_t)q * 3] - px, dy = spos[(size_t)q * 3 + 1] - py, dz = spos[(size_t)q * 3 + 2] - pz;
                  dsq = dx * dx + dy * dy + dz * dz;
                  keep = j != i && dsq < T;
                }
                const unsigned m = __ballot_sync(LQRO_FULL, keep);
                if (c + __popc(m) > LQRO_KNN_FAST_CAP) { fits = false; break; }
                if (keep) { const int pos = c + __popc(m & ((1u << lane) - 1u)); st.d[pos] = dsq; st.i[pos] = j; }
                c += __popc(m);
              }
            }
          __syncwarp();
          if (fits && c >= K) {
            // rank selection: entry e = lane + 32 t; its rank = number of staged keys below it
            double md[LQRO_KNN_FAST_CAP / 32];
            int mi[LQRO_KNN_FAST_CAP / 32], rk[LQRO_KNN_FAST_CAP / 32];
#pragma unroll
            for (int t = 0; t < LQRO_KNN_FAST_CAP / 32; ++t) {
              const int e = lane + 32 * t;
              md[t] = e < c ? st.d[e] : INF; mi[t] = e < c ? st.i[e] : 0x7fffffff; rk[t] = 0;
            }
            for (int u = 0; u < c; ++u) {
              const double od = st.d[u];
              const int oi = st.i[u];
#pragma unroll
              for (int t = 0; t < LQRO_KNN_FAST_CAP / 32; ++t) rk[t] += knn_less(od, oi, md[t], mi[t]) ? 1 : 0;
            }
#pragma unroll
            for (int t = 0; t < LQRO_KNN_FAST_CAP / 32; ++t)
              if (lane + 32 * t < c && rk[t] < K) nbr_out[(size_t)il * K + rk[t]] = mi[t];
            done = true;
          }
          __syncwarp();
        }
      }
      if (done) continue;
    }
    // ---- general path: shells of cells of growing radius
    KnnTop<E> top;
#pragma unroll
    for (int j = 0; j < E; ++j) { top.d[j] = INF; top.i[j] = 0x7fffffff; }
    int nst = 0, seen = 0;  // seen: agents in the cells searched so far (self included)
    // K-th best so far (entry K-1 of the sorted top), warp-uniform
    double kd = INF;
    int ki = 0x7fffffff;
    const int rmax = max(max(g.dim[0], g.dim[1]), g.dim[2]);
    for (int r = 0; r <= rmax; ++r) {
      const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dim[2] - 1), y0 = max(cy - r, 0), y1 = min(cy + r, g.dim[1] - 1);
      const int xa = cx - r, xb = cx + r;
      for (int zz = z0; zz <= z1; ++zz) {
        for (int yy = y0; yy <= y1; ++yy) {
          const bool face = (abs(zz - cz) == r) || (abs(yy - cy) == r);
          const int rowbase = (zz * g.dim[1] + yy) * g.dim[0];
          // face rows: the whole x range; inner rows: only the two end cells of the shell
          for (int part = 0; part < (face || r == 0 ? 1 : 2); ++part) {
            int c0, c1;
            if (face || r == 0) { c0 = max(xa, 0); c1 = min(xb, g.dim[0] - 1); }
            else { c0 = c1 = part == 0 ? xa : xb; if (c0 < 0 || c0 >= g.dim[0]) continue; }
            const int s0 = start[rowbase + c0], s1 = start[rowbase + c1 + 1];
            seen += s1 - s0;
            for (int b0 = s0; b0 < s1; b0 += 32) {
              const int q = b0 + lane;
              bool keep = false;
              double dsq = 0;
              int j = 0;
              if (q < s1) {
                j = sid[q];
                const double dx = spos[(size_t)q * 3] - px, dy = spos[(size_t)q * 3 + 1] - py, dz = spos[(size_t)q * 3 + 2] - pz;
                dsq = dx * dx + dy * dy + dz * dz;
                keep = j != i && knn_less(dsq, j, kd, ki);
              }
              const unsigned m = __ballot_sync(LQRO_FULL, keep);
              if (keep) { const int pos = nst + __popc(m & ((1u << lane) - 1u)); st.d[pos] = dsq; st.i[pos] = j; }
              nst += __popc(m);
              __syncwarp();
              if (nst >= 32 * E) {
                knn_flush<E>(top, st, nst, lane);
                // refresh the K-th best
                const int ke = K - 1;
                double vd = INF; int vi = 0x7fffffff;
#pragma unroll
                for (int jj = 0; jj < E; ++jj) if ((ke >> 5) == jj) { vd = top.d[jj]; vi = top.i[jj]; }
                kd = __shfl_sync(LQRO_FULL, vd, ke & 31);
                ki = __shfl_sync(LQRO_FULL, vi, ke & 31);
              }
            }
          }
        }
      }
      // shell r done: merge what is staged, then test whether unsearched cells can still matter
      if (seen <= K && r < rmax) continue;  // fewer than K others so far: the test below cannot succeed
      while (nst > 0) knn_flush<E>(top, st, nst, lane);
      {
        const int ke = K - 1;
        double vd = INF; int vi = 0x7fffffff;
#pragma unroll
        for (int jj = 0; jj < E; ++jj) if ((ke >> 5) == jj) { vd = top.d[jj]; vi = top.i[jj]; }
        kd = __shfl_sync(LQRO_FULL, vd, ke & 31);
        ki = __shfl_sync(LQRO_FULL, vi, ke & 31);
      }
      // distance from p to the nearest face of the searched block that still has cells behind it
      double cov = INF;
      if (cx - r > 0) cov = fmin(cov, px - (g.org[0] + (double)(cx - r) * g.h));
      if (cx + r < g.dim[0] - 1) cov = fmin(cov, (g.org[0] + (double)(cx + r + 1) * g.h) - px);
      if (cy - r > 0) cov = fmin(cov, py - (g.org[1] + (double)(cy - r) * g.h));
      if (cy + r < g.dim[1] - 1) cov = fmin(cov, (g.org[1] + (double)(cy + r + 1) * g.h) - py);
      if (cz - r > 0) cov = fmin(cov, pz - (g.org[2] + (double)(cz - r) * g.h));
      if (cz + r < g.dim[2] - 1) cov = fmin(cov, (g.org[2] + (double)(cz + r + 1) * g.h) - pz);
      if (cov == INF) break;  // the block covers the whole grid
      if (cov > 0) {
        const double safe = cov * (1.0 - 1.0e-9) - 1.0e-9 * g.h;  // cell assignment is exact only up to rounding at the faces
        if (safe > 0 && kd < safe * safe) break;
      }
    }
    // write the list: entries 0..K-1 of the sorted top (index 0x7fffffff = none)
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int e = j * 32 + lane;
      if (e < K) nbr_out[(size_t)il * K + e] = top.i[j] == 0x7fffffff ? -1 : top.i[j];
    }
    __syncwarp();
  }
}
