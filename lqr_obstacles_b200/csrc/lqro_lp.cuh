// lqro_lp.cuh -- the per-agent 3-D linear program (calculateNewV MAIN:1101-1112 ->
// linearProgram3/2/1 MAIN:872-1033, fallback linearProgram4 MAIN:1036-1082; the RVO2-3D programs of
// van den Berg et al.).  FP32 throughout, like the reference (V3:340 `float val_[3]`,
// RVO_EPSILON = 1e-5f MAIN:841); compiled without FMA contraction so every branch matches the CPU.
//
// Mapping: one thread per agent.  The planes of an agent are compacted (slot order = the reference's
// j order) into planes_c[agent][0..n); linearProgram4's projected planes use proj[agent][0..K).
#pragma once
#include "lqro_common.cuh"

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ float dotf(const F3& a, const F3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // V3:181-184
__device__ __forceinline__ F3 addf(const F3& a, const F3& b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 subf(const F3& a, const F3& b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 mulf(const F3& a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 divf(const F3& a, float s) { const float inv = 1.0f / s; return f3(a.x * inv, a.y * inv, a.z * inv); }  // V3:206-211
__device__ __forceinline__ F3 crossf(const F3& a, const F3& b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float abssqf(const F3& a) { return dotf(a, a); }
__device__ __forceinline__ F3 normalizef(const F3& a) { return divf(a, sqrtf(dotf(a, a))); }
__device__ __forceinline__ float sqrf(float s) { return s * s; }
#define LQRO_RVO_EPS 0.00001f

__device__ __forceinline__ F3 pl_point(const float* pl, int i) { return f3(pl[6 * i], pl[6 * i + 1], pl[6 * i + 2]); }
__device__ __forceinline__ F3 pl_normal(const float* pl, int i) { return f3(pl[6 * i + 3], pl[6 * i + 4], pl[6 * i + 5]); }

// linearProgram1, MAIN:872-939
__device__ inline bool lp1(const float* pl, int planeNo, const F3& lpoint, const F3& ldir, float radius, const F3& opt, bool dirOpt,
                           F3& result) {
  const float dp = dotf(lpoint, ldir);
  const float disc = sqrf(dp) + sqrf(radius) - abssqf(lpoint);
  if (disc < 0.0f) return false;
  const float sq = sqrtf(disc);
  float tLeft = -dp - sq, tRight = -dp + sq;
  for (int i = 0; i < planeNo; ++i) {
    const F3 pi = pl_point(pl, i), ni = pl_normal(pl, i);
    const float numerator = dotf(subf(pi, lpoint), ni);
    const float denominator = dotf(ldir, ni);
    if (sqrf(denominator) <= LQRO_RVO_EPS) {
      if (numerator > 0.0f) return false;
      continue;
    }
    const float t = numerator / denominator;
    if (denominator >= 0.0f) tLeft = (tLeft > t) ? tLeft : t;
    else tRight = (tRight < t) ? tRight : t;
    if (tLeft > tRight) return false;
  }
  if (dirOpt) {
    if (dotf(opt, ldir) > 0.0f) result = addf(lpoint, mulf(ldir, tRight));
    else result = addf(lpoint, mulf(ldir, tLeft));
  } else {
    const float t = dotf(ldir, subf(opt, lpoint));
    if (t < tLeft) result = addf(lpoint, mulf(ldir, tLeft));
    else if (t > tRight) result = addf(lpoint, mulf(ldir, tRight));
    else result = addf(lpoint, mulf(ldir, t));
  }
  return true;
}

// linearProgram2, MAIN:942-1001
__device__ inline bool lp2(const float* pl, int planeNo, float radius, const F3& opt, bool dirOpt, F3& result) {
  const F3 pN = pl_point(pl, planeNo), nN = pl_normal(pl, planeNo);
  const float planeDist = dotf(pN, nN);
  const float planeDistSq = sqrf(planeDist);
  const float radiusSq = sqrf(radius);
  if (planeDistSq > radiusSq) return false;
  const float planeRadiusSq = radiusSq - planeDistSq;
  const F3 planeCenter = mulf(nN, planeDist);
  if (dirOpt) {
    const F3 pov = subf(opt, mulf(nN, dotf(opt, nN)));
    const float lsq = abssqf(pov);
    if (lsq <= LQRO_RVO_EPS) result = planeCenter;
    else result = addf(planeCenter, mulf(pov, sqrtf(planeRadiusSq / lsq)));
  } else {
    result = addf(opt, mulf(nN, dotf(subf(pN, opt), nN)));
    if (abssqf(result) > radiusSq) {
      const F3 pr = subf(result, planeCenter);
      const float prl = abssqf(pr);
      result = addf(planeCenter, mulf(pr, sqrtf(planeRadiusSq / prl)));
    }
  }
  for (int i = 0; i < planeNo; ++i) {
    const F3 pi = pl_point(pl, i), ni = pl_normal(pl, i);
    if (dotf(ni, subf(pi, result)) > 0.0f) {
      const F3 cp = crossf(ni, nN);
      if (abssqf(cp) <= LQRO_RVO_EPS) return false;
      const F3 ldir = normalizef(cp);
      const F3 lnormal = crossf(ldir, nN);
      const F3 lpoint = addf(pN, mulf(lnormal, dotf(subf(pi, pN), ni) / dotf(lnormal, ni)));
      if (!lp1(pl, i, lpoint, ldir, radius, opt, dirOpt, result)) return false;
    }
  }
  return true;
}

// linearProgram3, MAIN:1004-1033
__device__ inline int lp3(const float* pl, int n, float radius, const F3& opt, bool dirOpt, F3& result) {
  if (dirOpt) result = mulf(opt, radius);
  else if (abssqf(opt) > sqrf(radius)) result = mulf(normalizef(opt), radius);
  else result = opt;
  for (int i = 0; i < n; ++i) {
    if (dotf(pl_normal(pl, i), subf(pl_point(pl, i), result)) > 0.0f) {
      const F3 temp = result;
      if (!lp2(pl, i, radius, opt, dirOpt, result)) { result = temp; return i; }
    }
  }
  return n;
}

// linearProgram4, MAIN:1036-1082
__device__ inline void lp4(const float* pl, int n, int beginPlane, float radius, float* proj, F3& result) {
  float distance = 0.0f;
  for (int i = beginPlane; i < n; ++i) {
    const F3 pi = pl_point(pl, i), ni = pl_normal(pl, i);
    if (dotf(ni, subf(pi, result)) > distance) {
      int np = 0;
      for (int j = 0; j < i; ++j) {
        const F3 pj = pl_point(pl, j), nj = pl_normal(pl, j);
        F3 ppoint;
        const F3 cp = crossf(nj, ni);
        if (abssqf(cp) <= LQRO_RVO_EPS) {
          if (dotf(ni, nj) > 0.0f) continue;
          ppoint = mulf(addf(pi, pj), 0.5f);
        } else {
          const F3 lnormal = crossf(cp, ni);
          ppoint = addf(pi, mulf(lnormal, dotf(subf(pj, pi), nj) / dotf(lnormal, nj)));
        }
        const F3 pnormal = normalizef(subf(nj, ni));
        float* o = proj + 6 * np;
        o[0] = ppoint.x; o[1] = ppoint.y; o[2] = ppoint.z; o[3] = pnormal.x; o[4] = pnormal.y; o[5] = pnormal.z;
        ++np;
      }
      const F3 temp = result;
      if (lp3(proj, np, radius, ni, true, result) < np) result = temp;
      distance = dotf(ni, subf(pi, result));
    }
  }
}

// calculateNewV for `batch` agents.  planes: [batch][stride][6]; status (optional) selects the slots
// that carry a plane (compaction in slot order); counts (optional, when status == NULL) gives the
// number of leading valid planes.  planes_c / proj: scratch [batch][stride][6].
__global__ void __launch_bounds__(128) k_new_velocity(int batch, int stride, const float* __restrict__ planes,
                                                      const int* __restrict__ status, const int* __restrict__ counts,
                                                      const double* __restrict__ goal, float radius, float* __restrict__ planes_c,
                                                      float* __restrict__ proj, double* __restrict__ newV, int* __restrict__ nplanes_out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= batch) return;
  const float* src = planes + (size_t)a * stride * 6;
  float* pc = planes_c + (size_t)a * stride * 6;
  int n = 0;
  if (status) {
    // compaction in slot order.  The statuses of a window of 64 slots are read first (independent, vectorised loads)
    // into a bit mask, then the planes of the set bits are copied four at a time, so that the thread waits for a
    // handful of memory round trips instead of one per slot.
    const int* stp = status + (size_t)a * stride;
    for (int w0 = 0; w0 < stride; w0 += 64) {
      const int wn = stride - w0 < 64 ? stride - w0 : 64;
      unsigned long long valid = 0ull;
      if ((stride & 3) == 0) {
        for (int s = 0; s < wn; s += 4) {
          const int4 q = *reinterpret_cast<const int4*>(stp + w0 + s);
          valid |= (unsigned long long)(((q.x & 0xff) != 0) | (((q.y & 0xff) != 0) << 1) | (((q.z & 0xff) != 0) << 2) | (((q.w & 0xff) != 0) << 3)) << s;
        }
      } else {
        for (int s = 0; s < wn; ++s) valid |= (unsigned long long)((stp[w0 + s] & 0xff) != 0) << s;
      }
      while (valid) {
        int sl[4];
        float2 v[4][3];
        int m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          sl[q] = valid ? __ffsll((long long)valid) - 1 : -1;
          if (valid) { valid &= valid - 1; ++m; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (sl[q] >= 0) {
            const float2* sp = reinterpret_cast<const float2*>(src + 6 * (size_t)(w0 + sl[q]));
            v[q][0] = sp[0]; v[q][1] = sp[1]; v[q][2] = sp[2];
          }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (sl[q] >= 0) {
            float2* dp = reinterpret_cast<float2*>(pc + 6 * (size_t)(n + q));
            dp[0] = v[q][0]; dp[1] = v[q][1]; dp[2] = v[q][2];
          }
        n += m;
      }
    }
  } else {
    n = counts[a];
    for (int s = 0; s < n; ++s)
#pragma unroll
      for (int d = 0; d < 6; ++d) pc[6 * s + d] = src[6 * s + d];
  }
  // Vector3 prefVelocity_(goalVel[0..2]): FP64 -> FP32 (MAIN:1103)
  const F3 pref = f3((float)goal[3 * a], (float)goal[3 * a + 1], (float)goal[3 * a + 2]);
  F3 res = f3(0.f, 0.f, 0.f);
  const int fail = lp3(pc, n, radius, pref, false, res);                      // MAIN:1106
  if (fail < n) lp4(pc, n, fail, radius, proj + (size_t)a * stride * 6, res);  // MAIN:1107-1108
  newV[3 * a] = (double)res.x; newV[3 * a + 1] = (double)res.y; newV[3 * a + 2] = (double)res.z;  // MAIN:1110
  if (nplanes_out) nplanes_out[a] = n;
}

// ------------------------------------------------------------------------------------------------
// Warp-per-agent variant for small agent counts (the reference's own 4..64-quad scenarios: dense, most planes
// violated, linearProgram4 on most steps -- thread-per-agent leaves the GPU empty and runs O(K^3) serial FP32 code).
// Same programs, same per-plane arithmetic; what is spread over the lanes is order-independent:
//   linearProgram1: tLeft = max, tRight = min over the planes (exact under any order), "return false" on any plane;
//   linearProgram2 / 3 / 4: the scan for the NEXT violated plane (the planes before it leave `result` untouched);
//   linearProgram4: the projected planes of one step (independent per j, compacted in j order).
// Planes live in shared memory; `result` is uniform across the warp.
__device__ __forceinline__ unsigned f32_key(float v) { const unsigned u = __float_as_uint(v); return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float f32_unkey(unsigned k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }

__device__ __forceinline__ bool wlp1(const float* pl, int planeNo, const F3& lpoint, const F3& ldir, float radius, const F3& opt, bool dirOpt,
                                     F3& result, int lane) {
  const float dp = dotf(lpoint, ldir);
  const float disc = sqrf(dp) + sqrf(radius) - abssqf(lpoint);
  if (disc < 0.0f) return false;
  const float sq = sqrtf(disc);
  float tLeft = -dp - sq, tRight = -dp + sq;
  for (int c0 = 0; c0 < planeNo; c0 += 32) {
    const int i = c0 + lane;
    bool bad = false;
    unsigned kl = 0u, kr = 0xffffffffu;  // neutral elements of max / min on keys
    if (i < planeNo) {
      const F3 pi = pl_point(pl, i), ni = pl_normal(pl, i);
      const float numerator = dotf(subf(pi, lpoint), ni);
      const float denominator = dotf(ldir, ni);
      if (sqrf(denominator) <= LQRO_RVO_EPS) bad = numerator > 0.0f;
      else {
        const float t = numerator / denominator;
        if (denominator >= 0.0f) kl = f32_key(t); else kr = f32_key(t);
      }
    }
    if (__any_sync(LQRO_FULL, bad)) return false;
    const unsigned ml = __reduce_max_sync(LQRO_FULL, kl), mr = __reduce_min_sync(LQRO_FULL, kr);
    if (ml != 0u) { const float t = f32_unkey(ml); tLeft = (tLeft > t) ? tLeft : t; }
    if (mr != 0xffffffffu) { const float t = f32_unkey(mr); tRight = (tRight < t) ? tRight : t; }
    if (tLeft > tRight) return false;
  }
  if (dirOpt) {
    if (dotf(opt, ldir) > 0.0f) result = addf(lpoint, mulf(ldir, tRight));
    else result = addf(lpoint, mulf(ldir, tLeft));
  } else {
    const float t = dotf(ldir, subf(opt, lpoint));
    if (t < tLeft) result = addf(lpoint, mulf(ldir, tLeft));
    else if (t > tRight) result = addf(lpoint, mulf(ldir, tRight));
    else result = addf(lpoint, mulf(ldir, t));
  }
  return true;
}

// index of the first plane in [from, n) with normal.(point - result) > bound, or n
__device__ __forceinline__ int wlp_next_violated(const float* pl, int from, int n, const F3& result, float bound, int lane) {
  for (int c0 = from; c0 < n; c0 += 32) {
    const int i = c0 + lane;
    const bool v = i < n && dotf(pl_normal(pl, i), subf(pl_point(pl, i), result)) > bound;
    const unsigned m = __ballot_sync(LQRO_FULL, v);
    if (m) return c0 + __ffs(m) - 1;
  }
  return n;
}

__device__ __forceinline__ bool wlp2(const float* pl, int planeNo, float radius, const F3& opt, bool dirOpt, F3& result, int lane) {
  const F3 pN = pl_point(pl, planeNo), nN = pl_normal(pl, planeNo);
  const float planeDist = dotf(pN, nN);
  const float planeDistSq = sqrf(planeDist);
  const float radiusSq = sqrf(radius);
  if (planeDistSq > radiusSq) return false;
  const float planeRadiusSq = radiusSq - planeDistSq;
  const F3 planeCenter = mulf(nN, planeDist);
  if (dirOpt) {
    const F3 pov = subf(opt, mulf(nN, dotf(opt, nN)));
    const float lsq = abssqf(pov);
    if (lsq <= LQRO_RVO_EPS) result = planeCenter;
    else result = addf(planeCenter, mulf(pov, sqrtf(planeRadiusSq / lsq)));
  } else {
    result = addf(opt, mulf(nN, dotf(subf(pN, opt), nN)));
    if (abssqf(result) > radiusSq) {
      const F3 pr = subf(result, planeCenter);
      const float prl = abssqf(pr);
      result = addf(planeCenter, mulf(pr, sqrtf(planeRadiusSq / prl)));
    }
  }
  int i = 0;
  while ((i = wlp_next_violated(pl, i, planeNo, result, 0.0f, lane)) < planeNo) {
    const F3 pi = pl_point(pl, i), ni = pl_normal(pl, i);
    const F3 cp = crossf(ni, nN);
    if (abssqf(cp) <= LQRO_RVO_EPS) return false;
    const F3 ldir = normalizef(cp);
    const F3 lnormal = crossf(ldir, nN);
    const F3 lpoint = addf(pN, mulf(lnormal, dotf(subf(pi, pN), ni) / dotf(lnormal, ni)));
    if (!wlp1(pl, i, lpoint, ldir, radius, opt, dirOpt, result, lane)) return false;
    ++i;
  }
  return true;
}

__device__ __forceinline__ int wlp3(const float* pl, int n, float radius, const F3& opt, bool dirOpt, F3& result, int lane) {
  if (dirOpt) result = mulf(opt, radius);
  else if (abssqf(opt) > sqrf(radius)) result = mulf(normalizef(opt), radius);
  else result = opt;
  int i = 0;
  while ((i = wlp_next_violated(pl, i, n, result, 0.0f, lane)) < n) {
    const F3 temp = result;
    if (!wlp2(pl, i, radius, opt, dirOpt, result, lane)) { result = temp; return i; }
    ++i;
  }
  return n;
}

__device__ __forceinline__ void wlp4(const float* pl, int n, int beginPlane, float radius, float* proj, F3& result, int lane) {
  float distance = 0.0f;
  int i = beginPlane;
  while ((i = wlp_next_violated(pl, i, n, result, distance, lane)) < n) {
    const F3 pi = pl_point(pl, i), ni = pl_normal(pl, i);
    int np = 0;
    __syncwarp();
    for (int c0 = 0; c0 < i; c0 += 32) {
      const int j = c0 + lane;
      bool emit = false;
      F3 ppoint = f3(0.f, 0.f, 0.f), pnormal = f3(0.f, 0.f, 0.f);
      if (j < i) {
        const F3 pj = pl_point(pl, j), nj = pl_normal(pl, j);
        const F3 cp = crossf(nj, ni);
        emit = true;
        if (abssqf(cp) <= LQRO_RVO_EPS) {
          if (dotf(ni, nj) > 0.0f) emit = false;
          else ppoint = mulf(addf(pi, pj), 0.5f);
        } else {
          const F3 lnormal = crossf(cp, ni);
          ppoint = addf(pi, mulf(lnormal, dotf(subf(pj, pi), nj) / dotf(lnormal, nj)));
        }
        if (emit) pnormal = normalizef(subf(nj, ni));
      }
      const unsigned m = __ballot_sync(LQRO_FULL, emit);
      if (emit) {
        float* o = proj + 6 * (np + __popc(m & ((1u << lane) - 1u)));
        o[0] = ppoint.x; o[1] = ppoint.y; o[2] = ppoint.z; o[3] = pnormal.x; o[4] = pnormal.y; o[5] = pnormal.z;
      }
      np += __popc(m);
    }
    __syncwarp();
    const F3 temp = result;
    if (wlp3(proj, np, radius, ni, true, result, lane) < np) result = temp;
    distance = dotf(ni, subf(pi, result));
    ++i;
  }
}

// calculateNewV, one warp per agent; shared memory: 2 * stride * 6 floats per warp (planes, projected planes)
__global__ void __launch_bounds__(128) k_new_velocity_warp(int batch, int stride, const float* __restrict__ planes, const int* __restrict__ status,
                                                           const double* __restrict__ goal, float radius, double* __restrict__ newV,
                                                           int* __restrict__ nplanes_out) {
  extern __shared__ float lp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* pc = lp_smem + (size_t)warp * 12 * stride;
  float* proj = pc + 6 * (size_t)stride;
  for (int a = blockIdx.x * wpb + warp; a < batch; a += gridDim.x * wpb) {
    const float* src = planes + (size_t)a * stride * 6;
    int n = 0;
    __syncwarp();
    for (int s0 = 0; s0 < stride; s0 += 32) {  // compaction in slot order
      const int sl = s0 + lane;
      const bool v = sl < stride && (status[(size_t)a * stride + sl] & 0xff) != 0;
      const unsigned m = __ballot_sync(LQRO_FULL, v);
      if (v) {
        float* o = pc + 6 * (n + __popc(m & ((1u << lane) - 1u)));
        const float2* sp = reinterpret_cast<const float2*>(src + 6 * (size_t)sl);
        const float2 q0 = sp[0], q1 = sp[1], q2 = sp[2];
        o[0] = q0.x; o[1] = q0.y; o[2] = q1.x; o[3] = q1.y; o[4] = q2.x; o[5] = q2.y;
      }
      n += __popc(m);
    }
    __syncwarp();
    const F3 pref = f3((float)goal[3 * a], (float)goal[3 * a + 1], (float)goal[3 * a + 2]);  // MAIN:1103
    F3 res = f3(0.f, 0.f, 0.f);
    const int fail = wlp3(pc, n, radius, pref, false, res, lane);   // MAIN:1106
    if (fail < n) wlp4(pc, n, fail, radius, proj, res, lane);        // MAIN:1107-1108
    if (lane == 0) {
      newV[3 * a] = (double)res.x; newV[3 * a + 1] = (double)res.y; newV[3 * a + 2] = (double)res.z;  // MAIN:1110
      if (nplanes_out) nplanes_out[a] = n;
    }
  }
}

// RVO-form Agent::update (AGENT:176-180): velocity_ = newVelocity_; position_ += velocity_ * timeStep_.
__global__ void k_agent_update(int n3, float* __restrict__ pos, float* __restrict__ vel, const float* __restrict__ newVel, float dt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n3) return;
  const float v = newVel[t];
  vel[t] = v;
  pos[t] = pos[t] + v * dt;
}
