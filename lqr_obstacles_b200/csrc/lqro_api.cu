// lqro_api.cu -- context, launch logic and the extern "C" ABI declared in include/lqro.h.
// Single translation unit: the kernels live in the .cuh files included below.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/lqro.h"
#define LQRO_MAX_CLASSES 16
#define LQRO_NBR_CHUNKS 4  // measured at 65,536 agents x K=64: none 7.85 ms end to end, 4 -> 7.74, 8 -> 7.84
#include "lqro_common.cuh"
#include "lqro_gjk.cuh"
#include "lqro_knn.cuh"
#include "lqro_pair.cuh"
#include "lqro_hull.cuh"
#include "lqro_step.cuh"
#include "lqro_knn.cuh"
#include "lqro_lp.cuh"
#include "lqro_riccati.cuh"
#include "lqro_sim.cuh"
#include "lqro_p2p.cuh"

// ---- minimal NCCL surface, resolved at run time (no link-time dependency; the process may already
//      hold a libnccl, e.g. the one bundled with PyTorch) ------------------------------------------
typedef struct { char internal[128]; } lqro_ncclUniqueId;
typedef void* lqro_ncclComm_t;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(lqro_ncclUniqueId*) = nullptr;
  int (*CommInitRank)(lqro_ncclComm_t*, int, lqro_ncclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, lqro_ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(lqro_ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static bool load_nccl() {
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (h) break; }   // reuse a loaded one first
  if (!h) for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) return false;
  g_nccl.handle = h;
  g_nccl.GetUniqueId = (int (*)(lqro_ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(lqro_ncclComm_t*, int, lqro_ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, lqro_ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(lqro_ncclComm_t))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.CommDestroy;
  return g_nccl.ok;
}

// ---- context ---------------------------------------------------------------------------------------
struct lqro_ctx {
  lqro_config cfg;
  LqroParams prm;
  LqroTables tb;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool have_dynamics = false, stepped = false, have_nbr = false;  // have_nbr: a neighbour table is resident in d_nbr
  // tables / dynamics
  double *d_dyn = nullptr;   // A(256) B(64) L(64) E(12)
  double *d_NFt = nullptr, *d_NF = nullptr, *d_T = nullptr, *d_sph = nullptr, *d_CG = nullptr, *d_prn = nullptr;
  // heterogeneous dynamics: tables of the extra classes 1..15 (class 0 = the buffers above), the class
  // of every agent, the owned agents ordered by class, one batch counter per class
  struct ClassTab { double *NFt = nullptr, *NF = nullptr, *T = nullptr, *CG = nullptr, *prn = nullptr; bool set = false; };
  ClassTab cls[LQRO_MAX_CLASSES];
  LqroTables* d_tbs = nullptr;           // [LQRO_MAX_CLASSES] for the hull kernel
  int *d_agent_class = nullptr, *d_order = nullptr, *d_cls_ctr = nullptr;
  int cls_begin[LQRO_MAX_CLASSES] = {0}, cls_count[LQRO_MAX_CLASSES] = {0};
  bool hetero = false;
  // host-buffer steps of large swarms: the neighbour table is uploaded in LQRO_NBR_CHUNKS agent blocks on a copy stream while
  // the pair kernel of the previous block runs (one launch per block, each on its own stream, so block c+1 moves onto an SM
  // as soon as block c's CTA leaves it)
  const int* h_nbr_chunked = nullptr;
  cudaStream_t copy_stream = nullptr, chunk_stream[LQRO_NBR_CHUNKS] = {nullptr};
  cudaEvent_t ev_ready = nullptr, ev_up[LQRO_NBR_CHUNKS] = {nullptr}, ev_done[LQRO_NBR_CHUNKS] = {nullptr};
  int *d_identity = nullptr, *d_chunk_ctr = nullptr;
  // neighbour table of a step uploaded NEXT TO the pair kernel of the polling-hull step (launch_step): rows in two blocks on their
  // own stream, a mark word written by the copy engine behind each block, the pair kernel waits for the rows of a batch
  size_t nbr_stream_min_pairs = (size_t)1 << 21; bool last_nbr_streamed = false;
  const int* h_nbr_streamed = nullptr; const double* h_vpref_streamed = nullptr;
  cudaStream_t nbr_stream = nullptr; cudaEvent_t ev_nbr_go = nullptr, ev_nbr_up = nullptr;
  int *d_nbr_mark = nullptr, *h_nbr_marks = nullptr;
  // step buffers
  double *d_x = nullptr, *d_vpref = nullptr, *d_newV = nullptr, *d_dist = nullptr;
  int *d_nbr = nullptr, *d_status = nullptr, *d_nreach = nullptr, *d_nfacets = nullptr, *d_nplanes = nullptr;
  float *d_planes = nullptr, *d_planes_c = nullptr, *d_proj = nullptr;
  int* d_ctr = nullptr;       // [0] pair work counter [1] hull count [2] hull work counter [3] degenerate [4..7] spare
  int* d_hull_items = nullptr;
  char* d_hull_ws = nullptr;
  double* d_pair_ovf = nullptr; size_t pair_ovf_stride = 0;
  long long* d_prof = nullptr;
  // device neighbour search (lqro_knn.cuh), allocated on first use
  unsigned long long* d_knn_bb = nullptr;
  KnnGrid* d_knn_grid = nullptr;
  int *d_knn_cell = nullptr, *d_knn_count = nullptr, *d_knn_start = nullptr, *d_knn_sid = nullptr;
  double* d_knn_spos = nullptr;
  int device_knn = 0;
  size_t hull_ws_stride = 0;
  int hull_slots = 0, hull_cap = 0, cap_pts = 0;
  // launch geometry
  int pair_warps = 0, pair_blocks = 0, fused_hull = 0; size_t pair_smem = 0;
  int hull_wpb = 2, hull_blocks = 0, hull_per_sm = 1, hull_poll_warps = 1; size_t hull_smem = 0;
  // step of large swarms: SMs of the polling hull launch (the pair kernel's grid is that much smaller), its stream and the events that order it
  int overlap_sms = 0; cudaStream_t hull_stream = nullptr; cudaEvent_t ev_poll_go = nullptr, ev_poll1 = nullptr, ev_poll0 = nullptr; bool last_polling = false;
  // stage scratch (grown on demand)
  char* d_scratch = nullptr; size_t scratch_bytes = 0;
  char* d_stage_ws = nullptr; int stage_ws_pts = 0;
  // multi-GPU
  lqro_ncclComm_t comm = nullptr; int rank = 0, world = 1;
  // all-gather over peer memory (lqro_p2p.cuh): the peers' state buffers and flag words mapped through CUDA IPC
  bool p2p = false; P2PPeers peers; int* d_p2p_flags = nullptr; int p2p_epoch = 0, p2p_done_epoch = 0; int p2p_blocks = 32;
  // simulator-form agent update (lqro_sim.cuh)
  LqroSimModel* d_sim_model = nullptr; bool have_sim_model = false, sim_attr_set = false; int sim_m_diag = 0;
  char* d_sim_buf = nullptr; size_t sim_buf_bytes = 0;
  cudaEvent_t ev_sim[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_ric[2] = {nullptr, nullptr};  // around the last k_riccati launch
  char* d_sim_scratch = nullptr; size_t sim_scratch_bytes = 0;
  long long launches = 0;      // kernel launches of the last avoidance step (launch_step)
  long long sim_launches = 0;  // kernel launches of the simulator-form entry points, cumulative
  std::string err;
};

static std::string g_create_err;

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      char b__[512];                                                                                  \
      snprintf(b__, sizeof b__, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      if (ctx) ctx->err = b__; else g_create_err = b__;                                               \
      return LQRO_E_CUDA;                                                                             \
    }                                                                                                 \
  } while (0)

static int fail(lqro_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg; else g_create_err = msg;
  return code;
}

extern "C" const char* lqro_version(void) { return "lqro 0.1 (sm_100a)"; }

extern "C" void lqro_default_config(lqro_config* c) {
  memset(c, 0, sizeof *c);
  c->device = 0;
  c->n_agents = 6;          // NUM_QUADS, MAIN:30
  c->local_begin = 0;
  c->n_local = 0;
  c->max_neighbors = 5;
  c->horizon = 30;          // OBSTACLE_STEPS, MAIN:33
  c->n_samples = 50;        // NUM_POINTS, MAIN:34
  c->xy_radius = 0.26;      // MAIN:35
  c->z_radius = 0.75;       // MAIN:36
  c->max_speed = 30.0;      // MAIN:1367
  c->lp_radius = 100.0;     // MAIN:1102
  c->hull_input = LQRO_HULL_TEXT6;
  c->fix_facet0_normal = 1;
  c->max_hull_pairs = 0;
  c->hull_slots = 0;
}

extern "C" const char* lqro_last_error(const lqro_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int ensure_scratch(lqro_ctx* ctx, size_t bytes) {
  if (ctx->scratch_bytes >= bytes) return LQRO_OK;
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
  CU(cudaMalloc(&ctx->d_scratch, bytes));
  ctx->scratch_bytes = bytes;
  return LQRO_OK;
}

extern "C" int lqro_destroy(lqro_ctx* ctx) {
  if (!ctx) return LQRO_OK;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->p2p) {
    for (int r = 0; r < ctx->world; ++r)
      if (r != ctx->rank) { if (ctx->peers.x[r]) cudaIpcCloseMemHandle(ctx->peers.x[r]); if (ctx->peers.flags[r]) cudaIpcCloseMemHandle(ctx->peers.flags[r]); }
  }
  if (ctx->d_p2p_flags) cudaFree(ctx->d_p2p_flags);
  if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
  void* ptrs[] = {ctx->d_dyn, ctx->d_NFt, ctx->d_NF, ctx->d_T, ctx->d_sph, ctx->d_CG, ctx->d_prn, ctx->d_x, ctx->d_vpref, ctx->d_newV, ctx->d_dist, ctx->d_nbr,
                  ctx->d_status, ctx->d_nreach, ctx->d_nfacets, ctx->d_nplanes, ctx->d_planes, ctx->d_planes_c, ctx->d_proj, ctx->d_ctr,
                  ctx->d_hull_items, ctx->d_hull_ws, ctx->d_scratch, ctx->d_stage_ws, ctx->d_pair_ovf, ctx->d_prof,
                  ctx->d_knn_bb, ctx->d_knn_grid, ctx->d_knn_cell, ctx->d_knn_count, ctx->d_knn_start, ctx->d_knn_sid, ctx->d_knn_spos};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int c = 1; c < LQRO_MAX_CLASSES; ++c) {
    void* q[] = {ctx->cls[c].NFt, ctx->cls[c].NF, ctx->cls[c].T, ctx->cls[c].CG, ctx->cls[c].prn};
    for (void* p : q) if (p) cudaFree(p);
  }
  { void* q[] = {ctx->d_sim_model, ctx->d_sim_buf, ctx->d_sim_scratch}; for (void* p : q) if (p) cudaFree(p); }
  for (cudaEvent_t e : ctx->ev_sim) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_ric) if (e) cudaEventDestroy(e);
  { void* q[] = {ctx->d_tbs, ctx->d_agent_class, ctx->d_order, ctx->d_cls_ctr, ctx->d_identity, ctx->d_chunk_ctr}; for (void* p : q) if (p) cudaFree(p); }
  if (ctx->hull_stream) cudaStreamDestroy(ctx->hull_stream);
  if (ctx->ev_poll_go) cudaEventDestroy(ctx->ev_poll_go);
  if (ctx->ev_poll1) cudaEventDestroy(ctx->ev_poll1);
  if (ctx->ev_poll0) cudaEventDestroy(ctx->ev_poll0);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->nbr_stream) cudaStreamDestroy(ctx->nbr_stream);
  if (ctx->ev_nbr_go) cudaEventDestroy(ctx->ev_nbr_go);
  if (ctx->ev_nbr_up) cudaEventDestroy(ctx->ev_nbr_up);
  if (ctx->d_nbr_mark) cudaFree(ctx->d_nbr_mark);
  if (ctx->h_nbr_marks) cudaFreeHost(ctx->h_nbr_marks);
  if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
  for (int c = 0; c < LQRO_NBR_CHUNKS; ++c) {
    if (ctx->chunk_stream[c]) cudaStreamDestroy(ctx->chunk_stream[c]);
    if (ctx->ev_up[c]) cudaEventDestroy(ctx->ev_up[c]);
    if (ctx->ev_done[c]) cudaEventDestroy(ctx->ev_done[c]);
  }
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return LQRO_OK;
}

extern "C" int lqro_create(const lqro_config* cfg, lqro_ctx** out) {
  lqro_ctx* ctx = nullptr;
  if (!cfg || !out) return fail(nullptr, LQRO_E_ARG, "null argument");
  *out = nullptr;
  if (cfg->n_agents < 1 || cfg->max_neighbors < 1 || cfg->horizon < 1 || cfg->horizon > 64 || cfg->n_samples < 1 || cfg->horizon * cfg->n_samples > 8192 ||
      !(cfg->max_speed > 0) || cfg->local_begin < 0 || cfg->local_begin >= cfg->n_agents)
    return fail(nullptr, LQRO_E_ARG, "bad configuration");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, LQRO_E_CUDA, "no CUDA device: liblqro has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LQRO_E_ARG, "bad device ordinal");
  CU(cudaSetDevice(cfg->device));
  ctx = new lqro_ctx();
  ctx->cfg = *cfg;
  if (ctx->cfg.n_local <= 0) ctx->cfg.n_local = cfg->n_agents - cfg->local_begin;
  if (ctx->cfg.local_begin + ctx->cfg.n_local > cfg->n_agents) { delete ctx; return fail(nullptr, LQRO_E_ARG, "local block exceeds n_agents"); }
  const int N = cfg->n_agents, NL = ctx->cfg.n_local, K = cfg->max_neighbors, H = cfg->horizon, P = cfg->n_samples;
  // pair ids, queue entries and the enumeration of a launch (lqro_enum.cuh: each of <= 2 * 8 chunks padded to 32) are 32-bit
  if ((long long)NL * K + 2 * 8 * 32 >= 0x7fffffffLL) { delete ctx; return fail(nullptr, LQRO_E_LIMIT, "n_local * max_neighbors must stay below 2^31 pairs per rank"); }
  LqroParams& p = ctx->prm;
  p.N = N; p.local_begin = cfg->local_begin; p.n_local = NL; p.K = K; p.H = H; p.P = P; p.HP = H * P;
  p.max_speed = cfg->max_speed; p.r2 = cfg->max_speed * cfg->max_speed; p.inv_r2 = 1.0 / p.r2;
  p.r2lo = p.r2 * (1.0 - 1.0e-9); p.r2hi = p.r2 * (1.0 + 1.0e-9);
  p.lp_radius = (float)cfg->lp_radius; p.hull_input = cfg->hull_input; p.fix_facet0 = 1;
  p.prune = (cfg->flags & LQRO_FLAG_NO_PRUNE) ? 0 : 1;
  // width of the first chunk of the slot-major enumeration (the following ones double, lqro_step.cuh): the inside pairs -- 173 / 41 /
  // 10 / 1 / 0 ... of the 225 of the bench swarm sit in slots 0 / 1 / 2 / 3 / 4+ -- should reach the hull queue first
  p.seg_split = 0;
  p.first_slots = 1;
  if (const char* e = getenv("LQRO_FIRST_SLOTS")) p.first_slots = atoi(e);
  if (p.first_slots < 1) p.first_slots = 1;
  if (p.first_slots > 32) p.first_slots = 32;
  ctx->device_knn = (cfg->flags & LQRO_FLAG_DEVICE_KNN) ? 1 : 0;
  p.box[0] = p.box[1] = p.box[2] = 0.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { delete ctx; return fail(nullptr, LQRO_E_CUDA, "cudaGetDeviceProperties failed"); }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  auto bail = [&](int code) { lqro_destroy(ctx); return code; };
#define CUC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { g_create_err = std::string(#call) + ": " + cudaGetErrorString(e__); return bail(LQRO_E_CUDA); } } while (0)
  CUC(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  for (auto& e : ctx->ev) CUC(cudaEventCreate(&e));
  CUC(cudaMalloc(&ctx->d_dyn, sizeof(double) * (256 + 64 + 64 + 12)));
  CUC(cudaMalloc(&ctx->d_NFt, sizeof(double) * 48 * H));
  CUC(cudaMalloc(&ctx->d_NF, sizeof(double) * 48 * H));
  CUC(cudaMalloc(&ctx->d_T, sizeof(double) * 9 * H));
  CUC(cudaMalloc(&ctx->d_sph, sizeof(double) * 3 * P));
  CUC(cudaMalloc(&ctx->d_CG, sizeof(double) * 9 * H));
  CUC(cudaMalloc(&ctx->d_prn, sizeof(double) * LQRO_PRN * H));
  const size_t NP = (size_t)NL * K;
  CUC(cudaMalloc(&ctx->d_x, sizeof(double) * 16 * (size_t)N));
  CUC(cudaMalloc(&ctx->d_vpref, sizeof(double) * 3 * (size_t)NL));
  CUC(cudaMalloc(&ctx->d_newV, sizeof(double) * 3 * (size_t)NL));
  CUC(cudaMalloc(&ctx->d_nbr, sizeof(int) * NP));
  CUC(cudaMalloc(&ctx->d_status, sizeof(int) * NP));
  CUC(cudaMalloc(&ctx->d_nreach, sizeof(int) * NP));
  CUC(cudaMalloc(&ctx->d_nfacets, sizeof(int) * NP));
  CUC(cudaMalloc(&ctx->d_nplanes, sizeof(int) * (size_t)NL));
  CUC(cudaMalloc(&ctx->d_dist, sizeof(double) * NP));
  CUC(cudaMalloc(&ctx->d_planes, sizeof(float) * 6 * NP));
  CUC(cudaMalloc(&ctx->d_planes_c, sizeof(float) * 6 * NP));
  CUC(cudaMalloc(&ctx->d_proj, sizeof(float) * 6 * NP));
  CUC(cudaMalloc(&ctx->d_ctr, sizeof(int) * 16));
  CUC(cudaMalloc(&ctx->d_prof, sizeof(long long) * 32));
  CUC(cudaMemset(ctx->d_prof, 0, sizeof(long long) * 32));
  CUC(cudaMemset(ctx->d_ctr, 0, sizeof(int) * 16));
  CUC(cudaMemset(ctx->d_status, 0, sizeof(int) * NP));
  CUC(cudaMemset(ctx->d_nfacets, 0, sizeof(int) * NP));
  CUC(cudaMemset(ctx->d_nreach, 0, sizeof(int) * NP));
  CUC(cudaMemset(ctx->d_dist, 0, sizeof(double) * NP));
  CUC(cudaMemset(ctx->d_planes, 0, sizeof(float) * 6 * NP));
  ctx->hull_cap = cfg->max_hull_pairs > 0 ? cfg->max_hull_pairs : (int)NP;  // lqro.h: 0 => every pair may queue
  CUC(cudaMalloc(&ctx->d_hull_items, sizeof(int) * (size_t)ctx->hull_cap));
  CUC(cudaMemset(ctx->d_hull_items, 0xff, sizeof(int) * (size_t)ctx->hull_cap));  // -1 = not published; every consumer restores it (k_hull)
  // sphere samples: createSpheres, MAIN:665-680 (one-off table; radii doubled for the Minkowski sum)
  {
    std::vector<double> sph(3 * (size_t)P);
    const double dlong = M_PI * (3.0 - sqrt(5.0));
    const double dz = 2.0 / P;
    double longt = 0, z = 1.0 - dz / 2.0;
    for (int i = 0; i < P; ++i) {
      sph[3 * i + 0] = 2 * cfg->xy_radius * cos(longt) * sqrt(1 - z * z);
      sph[3 * i + 1] = 2 * cfg->xy_radius * sin(longt) * sqrt(1 - z * z);
      sph[3 * i + 2] = 2 * cfg->z_radius * z;
      z -= dz;
      longt += dlong;
    }
    CUC(cudaMemcpy(ctx->d_sph, sph.data(), sizeof(double) * 3 * P, cudaMemcpyHostToDevice));
    for (int i = 0; i < P; ++i)
      for (int k = 0; k < 3; ++k) if (fabs(sph[3 * i + k]) > p.box[k]) p.box[k] = fabs(sph[3 * i + k]);
  }
  ctx->tb.NFt = ctx->d_NFt; ctx->tb.T = ctx->d_T; ctx->tb.sph = ctx->d_sph; ctx->tb.CG = ctx->d_CG; ctx->tb.prn = ctx->d_prn;
  ctx->cls[0].NFt = ctx->d_NFt; ctx->cls[0].NF = ctx->d_NF; ctx->cls[0].T = ctx->d_T; ctx->cls[0].CG = ctx->d_CG; ctx->cls[0].prn = ctx->d_prn;
  // launch geometry of the pair kernel: as many warps per block as the opt-in shared memory holds
  {
    const size_t tbl = sizeof(double) * pair_block_table_doubles(H, P);
    const size_t per = sizeof(double) * pair_warp_doubles(H, P);     // full-capacity warp scratch (stage kernels)
    const size_t per2 = sizeof(double) * pair2_warp_doubles(H);      // pair kernel: capped cloud + overflow in HBM
    if (tbl + per > ctx->smem_optin) { g_create_err = "H*P too large for shared memory"; return bail(LQRO_E_LIMIT); }
    // one hull warp per CTA if its structures fit next to >= 8 pair warps (512 threads x 128 registers = the
    // whole register file, so 15 + 1 warps at most)
    const size_t hullw = hull_warp_smem_bytes(H, P) + 32;
    ctx->fused_hull = ((cfg->flags & LQRO_FLAG_FUSED_HULL) && tbl + 8 * per2 + hullw <= ctx->smem_optin) ? 1 : 0;
    int wp = (int)((ctx->smem_optin - tbl - (ctx->fused_hull ? hullw : 0)) / per2);
    const int wmax = ctx->fused_hull ? LQRO_PAIR_MAX_WARPS - 1 : LQRO_PAIR_MAX_WARPS;
    if (wp > wmax) wp = wmax;
    if (wp < 1) { g_create_err = "H too large for shared memory"; return bail(LQRO_E_LIMIT); }
    ctx->pair_warps = wp;
    ctx->pair_smem = tbl + per2 * wp + (ctx->fused_hull ? hullw : 0);
    ctx->pair_blocks = ctx->sm_count;  // persistent: one CTA per SM
    ctx->pair_ovf_stride = 3 * (size_t)(H * P > LQRO_PAIR_CLOUD_CAP ? H * P - LQRO_PAIR_CLOUD_CAP : 0) + 8;
    // one overflow area per pair warp AND per concurrent launch: the chunked neighbour upload (launch_step) runs LQRO_NBR_CHUNKS
    // pair kernels on independent streams, CTA b of chunk c and CTA b of chunk c + 1 can be resident at the same time
    CUC(cudaMalloc(&ctx->d_pair_ovf, sizeof(double) * ctx->pair_ovf_stride * (size_t)ctx->pair_blocks * wp * LQRO_NBR_CHUNKS));
    // function attributes are per function, not per context: always allow the device maximum so that
    // contexts with different H*P can coexist
    CUC(cudaFuncSetAttribute(k_pairs<false, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_pairs<false, 20, 50>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_pairs<false, 30, 50>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_pairs<true, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_pairs<false, 0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_pairs<false, 20, 50, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_pairs<false, 30, 50, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    CUC(cudaFuncSetAttribute(k_stage_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    ctx->hull_wpb = 1;
    ctx->cap_pts = H * P;
    ctx->hull_smem = hull_block_smem_bytes(H, P);
    if (ctx->hull_smem > ctx->smem_optin) { g_create_err = "H*P too large for the hull kernel's shared memory"; return bail(LQRO_E_LIMIT); }
    CUC(cudaFuncSetAttribute(k_hull, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    int per_sm = (int)((ctx->smem_optin + 1024) / (ctx->hull_smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    ctx->hull_per_sm = per_sm;
    int slots = cfg->hull_slots > 0 ? cfg->hull_slots : ctx->sm_count * per_sm;
    if (slots < ctx->sm_count) slots = ctx->sm_count;
    ctx->hull_slots = slots;
    ctx->hull_blocks = slots;
    // SMs of the polling hull launch of a large swarm's step (the pair kernel's grid is that much smaller); LQRO_HULL_OVERLAP_SMS=0
    // or LQRO_FLAG_NO_OVERLAP: pair kernel on every SM, hulls afterwards
    ctx->overlap_sms = (cfg->flags & LQRO_FLAG_NO_OVERLAP) ? 0 : 12;  // measured at 65,536 x 64: 8 -> 6.79 ms (hulls behind), 12 -> 6.40, 16 -> 6.59
    if (const char* e = getenv("LQRO_HULL_OVERLAP_SMS")) ctx->overlap_sms = atoi(e);
    if (const char* e = getenv("LQRO_NBR_STREAM_MIN_PAIRS")) { const long long v = atoll(e); ctx->nbr_stream_min_pairs = v > 0 ? (size_t)v : ~(size_t)0; }
    if (ctx->overlap_sms < 0) ctx->overlap_sms = 0;
    if (ctx->overlap_sms > ctx->sm_count / 2) ctx->overlap_sms = ctx->sm_count / 2;
    // hull warps per block of the polling launch: as many as the opt-in shared memory holds, at most LQRO_HULL_POLL_WARPS (its
    // shared-memory REQUEST is always the whole opt-in size, which keeps pair CTAs off its SMs)
    ctx->hull_poll_warps = (int)((ctx->smem_optin - hull_align16(sizeof(double) * 3 * (size_t)P)) / hull_warp_smem_bytes(H, P));
    if (ctx->hull_poll_warps > LQRO_HULL_POLL_WARPS) ctx->hull_poll_warps = LQRO_HULL_POLL_WARPS;
    if (ctx->hull_poll_warps < 1) ctx->hull_poll_warps = 1;
    if (ctx->overlap_sms * ctx->hull_poll_warps > slots) ctx->overlap_sms = slots / ctx->hull_poll_warps;
    ctx->hull_ws_stride = hull_ws_bytes(ctx->cap_pts);
    // global slots: one per hull of the wide launch, then one per hull of the polling launch (the two can be resident together)
    CUC(cudaMalloc(&ctx->d_hull_ws, ctx->hull_ws_stride * ((size_t)slots + (size_t)ctx->overlap_sms * ctx->hull_poll_warps)));
  }
  CUC(cudaFuncSetAttribute(k_riccati, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * ((sizeof(RiccatiWs) + 7) / 8) * 8)));
#undef CUC
  *out = ctx;
  return LQRO_OK;
}

// ---- (1) gains and tables ---------------------------------------------------------------------------
extern "C" int lqro_solve_gains(lqro_ctx* ctx, int nb, const double* A, const double* B, double qv, double qp, double r, int iters,
                                double* L, double* E, double* Lh, double* Eh) {
  if (!ctx || nb < 1 || !A || !B || !L || !E || !Lh || !Eh || iters < 1) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  const size_t nin = (size_t)nb * (256 + 64), nout = (size_t)nb * (64 + 12 + 48 + 9);
  int rc = ensure_scratch(ctx, sizeof(double) * (nin + nout));
  if (rc) return rc;
  double* dA = (double*)ctx->d_scratch; double* dB = dA + (size_t)nb * 256;
  double* dL = dB + (size_t)nb * 64; double* dE = dL + (size_t)nb * 64; double* dLh = dE + (size_t)nb * 12; double* dEh = dLh + (size_t)nb * 48;
  CU(cudaMemcpyAsync(dA, A, sizeof(double) * 256 * nb, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dB, B, sizeof(double) * 64 * nb, cudaMemcpyHostToDevice, ctx->stream));
  const int wpb = 4;
  const size_t smem = (size_t)wpb * ((sizeof(RiccatiWs) + 7) / 8) * 8;
  for (cudaEvent_t& e : ctx->ev_ric) if (!e) CU(cudaEventCreate(&e));
  CU(cudaEventRecord(ctx->ev_ric[0], ctx->stream));
  k_riccati<<<(nb + wpb - 1) / wpb, wpb * 32, smem, ctx->stream>>>(nb, dA, dB, qv, qp, r, iters, dL, dE, dLh, dEh);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev_ric[1], ctx->stream));
  CU(cudaMemcpyAsync(L, dL, sizeof(double) * 64 * nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(E, dE, sizeof(double) * 12 * nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(Lh, dLh, sizeof(double) * 48 * nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(Eh, dEh, sizeof(double) * 9 * nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_riccati_last_ms(lqro_ctx* ctx, float* ms) {
  if (!ctx || !ms) return LQRO_E_ARG;
  if (!ctx->ev_ric[0]) return fail(ctx, LQRO_E_STATE, "no lqro_solve_gains yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaEventSynchronize(ctx->ev_ric[1]));
  CU(cudaEventElapsedTime(ms, ctx->ev_ric[0], ctx->ev_ric[1]));
  return LQRO_OK;
}

static int set_dynamics_impl(lqro_ctx* ctx, int cls, const double* A, const double* B, const double* L, const double* E) {
  if (!ctx || !A || !B || !L || !E || cls < 0 || cls >= LQRO_MAX_CLASSES) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  lqro_ctx::ClassTab& ct = ctx->cls[cls];
  if (!ct.NFt) {
    const int H = ctx->prm.H;
    CU(cudaMalloc(&ct.NFt, sizeof(double) * 48 * H));
    CU(cudaMalloc(&ct.NF, sizeof(double) * 48 * H));
    CU(cudaMalloc(&ct.T, sizeof(double) * 9 * H));
    CU(cudaMalloc(&ct.CG, sizeof(double) * 9 * H));
    CU(cudaMalloc(&ct.prn, sizeof(double) * LQRO_PRN * H));
  }
  double* d = ctx->d_dyn;
  CU(cudaMemcpyAsync(d, A, sizeof(double) * 256, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d + 256, B, sizeof(double) * 64, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d + 320, L, sizeof(double) * 64, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d + 384, E, sizeof(double) * 12, cudaMemcpyHostToDevice, ctx->stream));
  k_horizon_tables<<<1, 256, 0, ctx->stream>>>(d, d + 256, d + 320, d + 384, ctx->prm.H, ct.NFt, ct.NF, ct.T, ct.CG);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  // pruning constants per horizon step (see step_prunable in lqro_pair.cuh): an UPPER bound of
  // sigma_max(C G_k) (largest eigenvalue of G'G by Jacobi sweeps, inflated by 1e-9; never below the
  // true value, falling back to the Frobenius norm) and kappa_k = |T_k|_F * sigma_max.
  {
    const int H = ctx->prm.H;
    std::vector<double> cg(9 * (size_t)H), T(9 * (size_t)H), prn(LQRO_PRN * (size_t)H, 0.0);
    CU(cudaMemcpy(cg.data(), ct.CG, sizeof(double) * 9 * H, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(T.data(), ct.T, sizeof(double) * 9 * H, cudaMemcpyDeviceToHost));
    std::vector<double> nf(48 * (size_t)H), nfmax(H, 0.0);
    CU(cudaMemcpy(nf.data(), ct.NF, sizeof(double) * 48 * H, cudaMemcpyDeviceToHost));
    for (int k = 0; k < H; ++k) for (int e = 0; e < 48; ++e) nfmax[k] = fmax(nfmax[k], fabs(nf[48 * (size_t)k + e]));
    for (int k = 0; k < H; ++k) {
      const double* g = cg.data() + 9 * k;
      double m[3][3];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { m[i][j] = 0; for (int q = 0; q < 3; ++q) m[i][j] += g[3 * q + i] * g[3 * q + j]; }
      double fro2 = m[0][0] + m[1][1] + m[2][2];
      for (int sweep = 0; sweep < 30; ++sweep) {  // cyclic Jacobi on the symmetric 3x3
        for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) {
          if (fabs(m[a][b]) < 1e-300) continue;
          const double th = 0.5 * atan2(2 * m[a][b], m[a][a] - m[b][b]);
          const double c = cos(th), s2 = sin(th);
          double r[3][3];
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = (i == j) ? 1.0 : 0.0;
          r[a][a] = c; r[b][b] = c; r[a][b] = -s2; r[b][a] = s2;
          double t1[3][3], t2[3][3];
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { t1[i][j] = 0; for (int q = 0; q < 3; ++q) t1[i][j] += m[i][q] * r[q][j]; }
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { t2[i][j] = 0; for (int q = 0; q < 3; ++q) t2[i][j] += r[q][i] * t1[q][j]; }
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = t2[i][j];
        }
      }
      double lam = fmax(m[0][0], fmax(m[1][1], m[2][2]));
      const double off = fabs(m[0][1]) + fabs(m[0][2]) + fabs(m[1][2]);
      lam += 2 * off;                                  // Gershgorin slack for any residual off-diagonal mass
      if (!(lam > 0) || !(lam <= fro2 * (1 + 1e-9))) lam = fro2;   // safe fallback: sigma_max <= |G|_F
      const double smax = sqrt(lam) * (1.0 + 1e-9);
      double tf = 0, ef = 0;
      for (int e = 0; e < 9; ++e) { tf += T[9 * k + e] * T[9 * k + e]; if (e % 4 != 0) ef += T[9 * k + e] * T[9 * k + e]; }
      double* c = prn.data() + LQRO_PRN * (size_t)k;
      c[0] = (ctx->prm.max_speed * smax) / ctx->prm.max_speed;  // the quotient the kernel used to form per step and lane (same IEEE division, same bits)
      c[1] = sqrt(tf) * smax * (1.0 + 1e-9);
      c[2] = fabs(T[9 * k + 0]); c[3] = fabs(T[9 * k + 4]); c[4] = fabs(T[9 * k + 8]);
      c[5] = sqrt(ef) * (1.0 + 1e-9);
      c[6] = sqrt(tf) * (1.0 + 1e-9);
      c[7] = nfmax[k];
    }
    CU(cudaMemcpy(ct.prn, prn.data(), sizeof(double) * LQRO_PRN * H, cudaMemcpyHostToDevice));
  }
  ct.set = true;
  if (cls == 0) ctx->have_dynamics = true;
  {
    LqroTables tbs[LQRO_MAX_CLASSES];
    for (int c = 0; c < LQRO_MAX_CLASSES; ++c) {
      const lqro_ctx::ClassTab& q = ctx->cls[c].set ? ctx->cls[c] : ctx->cls[0];
      tbs[c].NFt = q.NFt; tbs[c].T = q.T; tbs[c].sph = ctx->d_sph; tbs[c].CG = q.CG; tbs[c].prn = q.prn;
    }
    if (!ctx->d_tbs) CU(cudaMalloc(&ctx->d_tbs, sizeof(tbs)));
    CU(cudaMemcpy(ctx->d_tbs, tbs, sizeof(tbs), cudaMemcpyHostToDevice));
  }
  return LQRO_OK;
}

extern "C" int lqro_set_dynamics(lqro_ctx* ctx, const double* A, const double* B, const double* L, const double* E) {
  return set_dynamics_impl(ctx, 0, A, B, L, E);
}

extern "C" int lqro_set_dynamics_class(lqro_ctx* ctx, int cls, const double* A, const double* B, const double* L, const double* E) {
  return set_dynamics_impl(ctx, cls, A, B, L, E);
}

// class of every agent (n_agents ints in [0, 16), every class used must have its dynamics set); NULL = all class 0
extern "C" int lqro_set_agent_classes(lqro_ctx* ctx, const int* cls) {
  if (!ctx) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  const LqroParams& p = ctx->prm;
  if (!cls) { ctx->hetero = false; return LQRO_OK; }
  // validate into locals, commit only on success (a failed call leaves the previous assignment in force)
  int count[LQRO_MAX_CLASSES] = {0}, begin[LQRO_MAX_CLASSES] = {0};
  for (int i = 0; i < p.N; ++i) {
    if (cls[i] < 0 || cls[i] >= LQRO_MAX_CLASSES || !ctx->cls[cls[i]].set) return fail(ctx, LQRO_E_ARG, "agent class without dynamics");
  }
  for (int il = 0; il < p.n_local; ++il) count[cls[p.local_begin + il]]++;
  int run = 0;
  for (int c = 0; c < LQRO_MAX_CLASSES; ++c) { begin[c] = run; run += count[c]; }
  std::vector<int> order(p.n_local), fill(LQRO_MAX_CLASSES, 0);
  for (int il = 0; il < p.n_local; ++il) { const int c = cls[p.local_begin + il]; order[begin[c] + fill[c]++] = il; }
  if (!ctx->d_agent_class) {
    CU(cudaMalloc(&ctx->d_agent_class, sizeof(int) * (size_t)p.N));
    CU(cudaMalloc(&ctx->d_order, sizeof(int) * (size_t)p.n_local));
    CU(cudaMalloc(&ctx->d_cls_ctr, sizeof(int) * LQRO_MAX_CLASSES));
  }
  CU(cudaMemcpy(ctx->d_agent_class, cls, sizeof(int) * (size_t)p.N, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(ctx->d_order, order.data(), sizeof(int) * (size_t)p.n_local, cudaMemcpyHostToDevice));
  for (int c = 0; c < LQRO_MAX_CLASSES; ++c) { ctx->cls_count[c] = count[c]; ctx->cls_begin[c] = begin[c]; }
  ctx->hetero = true;
  return LQRO_OK;
}

extern "C" int lqro_get_spheres(lqro_ctx* ctx, double* sph) {
  if (!ctx || !sph) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaMemcpy(sph, ctx->d_sph, sizeof(double) * 3 * ctx->prm.P, cudaMemcpyDeviceToHost));
  return LQRO_OK;
}

extern "C" int lqro_get_tables(lqro_ctx* ctx, double* NF, double* T, double* sph) {
  if (!ctx) return LQRO_E_ARG;
  if (!ctx->have_dynamics) return fail(ctx, LQRO_E_STATE, "set_dynamics first");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  if (NF) CU(cudaMemcpy(NF, ctx->d_NF, sizeof(double) * 48 * ctx->prm.H, cudaMemcpyDeviceToHost));
  if (T) CU(cudaMemcpy(T, ctx->d_T, sizeof(double) * 9 * ctx->prm.H, cudaMemcpyDeviceToHost));
  if (sph) CU(cudaMemcpy(sph, ctx->d_sph, sizeof(double) * 3 * ctx->prm.P, cudaMemcpyDeviceToHost));
  return LQRO_OK;
}

// ---- neighbour search on the device (AGENT:153-174 ordering) ----------------------------
static int launch_knn(lqro_ctx* ctx, const double* d_x, int* d_nbr_out, cudaStream_t st) {
  const LqroParams& p = ctx->prm;
  if (p.K > 128) return fail(ctx, LQRO_E_LIMIT, "device neighbour search supports K <= 128");
  if (!ctx->d_knn_bb) {
    CU(cudaMalloc(&ctx->d_knn_bb, sizeof(unsigned long long) * 6));
    CU(cudaMalloc(&ctx->d_knn_grid, sizeof(KnnGrid)));
    CU(cudaMalloc(&ctx->d_knn_cell, sizeof(int) * (size_t)p.N));
    CU(cudaMalloc(&ctx->d_knn_count, sizeof(int) * LQRO_KNN_MAX_CELLS));
    CU(cudaMalloc(&ctx->d_knn_start, sizeof(int) * (LQRO_KNN_MAX_CELLS + 1)));
    CU(cudaMalloc(&ctx->d_knn_sid, sizeof(int) * (size_t)p.N));
    CU(cudaMalloc(&ctx->d_knn_spos, sizeof(double) * 3 * (size_t)p.N));
  }
  CU(cudaMemsetAsync(ctx->d_knn_bb, 0xff, sizeof(unsigned long long) * 3, st));
  CU(cudaMemsetAsync(ctx->d_knn_bb + 3, 0, sizeof(unsigned long long) * 3, st));
  CU(cudaMemsetAsync(ctx->d_knn_count, 0, sizeof(int) * LQRO_KNN_MAX_CELLS, st));
  const int nb = std::min((p.N + 255) / 256, 2 * ctx->sm_count);
  k_knn_bbox<<<nb, 256, 0, st>>>(d_x, p.N, ctx->d_knn_bb);
  k_knn_setup<<<1, 32, 0, st>>>(ctx->d_knn_bb, p.N, p.K, ctx->d_knn_grid);
  k_knn_count<<<nb, 256, 0, st>>>(d_x, p.N, ctx->d_knn_grid, ctx->d_knn_cell, ctx->d_knn_count);
  k_knn_scan<<<1, 1024, 0, st>>>(ctx->d_knn_grid, ctx->d_knn_count, ctx->d_knn_start);
  k_knn_scatter<<<nb, 256, 0, st>>>(d_x, p.N, ctx->d_knn_cell, ctx->d_knn_start, ctx->d_knn_count, ctx->d_knn_sid, ctx->d_knn_spos);
  const int E = p.K <= 32 ? 1 : p.K <= 64 ? 2 : 4;
  const int wpb = 4;
  const size_t smem = (size_t)wpb * knn_warp_smem_bytes(E);
  const int qb = std::min((p.n_local + wpb - 1) / wpb, 16 * ctx->sm_count);
  if (E == 1) k_knn_query<1><<<qb, wpb * 32, smem, st>>>(d_x, p.local_begin, p.n_local, p.K, ctx->d_knn_grid, ctx->d_knn_start, ctx->d_knn_sid, ctx->d_knn_spos, d_nbr_out);
  else if (E == 2) k_knn_query<2><<<qb, wpb * 32, smem, st>>>(d_x, p.local_begin, p.n_local, p.K, ctx->d_knn_grid, ctx->d_knn_start, ctx->d_knn_sid, ctx->d_knn_spos, d_nbr_out);
  else k_knn_query<4><<<qb, wpb * 32, smem, st>>>(d_x, p.local_begin, p.n_local, p.K, ctx->d_knn_grid, ctx->d_knn_start, ctx->d_knn_sid, ctx->d_knn_spos, d_nbr_out);
  CU(cudaGetLastError());
  if (d_nbr_out == ctx->d_nbr) ctx->have_nbr = true;
  return LQRO_OK;
}

extern "C" int lqro_knn_device(lqro_ctx* ctx, const double* d_x, int* d_nbr, void* stream) {
  if (!ctx) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  return launch_knn(ctx, d_x ? d_x : ctx->d_x, d_nbr ? d_nbr : ctx->d_nbr, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int lqro_knn(lqro_ctx* ctx, const double* x, int* nbr) {
  if (!ctx || !nbr) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  const LqroParams& p = ctx->prm;
  cudaStream_t st = ctx->stream;
  if (x) CU(cudaMemcpyAsync(ctx->d_x, x, sizeof(double) * 16 * (size_t)p.N, cudaMemcpyHostToDevice, st));
  int rc = launch_knn(ctx, ctx->d_x, ctx->d_nbr, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(nbr, ctx->d_nbr, sizeof(int) * (size_t)p.n_local * p.K, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LQRO_OK;
}

extern "C" int lqro_get_neighbors(lqro_ctx* ctx, int* nbr) {
  if (!ctx || !nbr) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(nbr, ctx->d_nbr, sizeof(int) * (size_t)ctx->prm.n_local * ctx->prm.K, cudaMemcpyDeviceToHost));
  return LQRO_OK;
}

// ---- the step ----------------------------------------------------------------------------------------
// batches of the nearest 8 slots in a polling-hull step: narrow enough for >= 6 batches per pair warp of the grid (a batch is one
// warp's sequential job)
static int polling_batch_width(const lqro_ctx* ctx) {
  const LqroParams& p = ctx->prm;
  const long long narrow = slot_major_prefix(p.n_local, p.K, p.first_slots, 8);
  const int pair_grid = ctx->pair_blocks - ctx->overlap_sms;
  int bw = 32;
  while (bw > 2 && narrow / bw < 6LL * pair_grid * ctx->pair_warps) bw >>= 1;
  return bw;
}
// Host neighbour table (and preferred velocities) of a host-buffer step: which way up?  Blocks next to several pair launches
// (no polling hull launch), rows streamed next to THE pair launch of a polling-hull step, or a plain copy in front of the step.
// Streaming starts at 2^21 pairs (LQRO_NBR_STREAM_MIN_PAIRS; 0 = never): the agents of the second block surface their inside pairs a
// twelfth of the pair kernel later, which a smaller rank -- whose polling hull launch already ends with the pair kernel -- would
// pay for with exposed hull time.
static int upload_step_inputs(lqro_ctx* ctx, const double* vpref, const int* nbr, cudaStream_t st) {
  const LqroParams& p = ctx->prm;
  const bool big = nbr && !ctx->hetero && !ctx->fused_hull && (size_t)p.n_local * p.K >= ((size_t)1 << 20);
  const bool chunked = big && ctx->overlap_sms == 0;
  const bool streamed = nbr && !ctx->hetero && !ctx->fused_hull && ctx->overlap_sms > 0 && p.n_local >= 2048 && (size_t)p.n_local * p.K >= ctx->nbr_stream_min_pairs &&
                        (size_t)p.n_local * p.K >= ((size_t)1 << 17);  // the polling-hull step of launch_step
  ctx->h_nbr_chunked = chunked ? nbr : nullptr;
  ctx->h_nbr_streamed = streamed ? nbr : nullptr;
  ctx->h_vpref_streamed = streamed ? vpref : nullptr;
  if (!streamed) CU(cudaMemcpyAsync(ctx->d_vpref, vpref, sizeof(double) * 3 * (size_t)p.n_local, cudaMemcpyHostToDevice, st));
  if (nbr && !chunked && !streamed) CU(cudaMemcpyAsync(ctx->d_nbr, nbr, sizeof(int) * (size_t)p.n_local * p.K, cudaMemcpyHostToDevice, st));
  return LQRO_OK;
}

static int launch_step(lqro_ctx* ctx, const double* d_x, const double* d_vpref, const int* d_nbr, double* d_newV, cudaStream_t st) {
  const LqroParams& p = ctx->prm;
  PairOutputs po{ctx->d_planes, ctx->d_status, ctx->d_nreach, ctx->d_dist, ctx->d_nfacets};
  HullQueue hq{ctx->d_hull_items, ctx->d_ctr + 1, ctx->hull_cap};
  CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(int) * 16, st));
  CU(cudaMemsetAsync(ctx->d_prof, 0, sizeof(long long) * 32, st));
  if (ctx->fused_hull) CU(cudaMemsetAsync(ctx->d_hull_items, 0xff, sizeof(int) * (size_t)ctx->hull_cap, st));  // -1: ticket not filled yet (polled by the fused hull warps only)
  // d_ctr: [0] batch counter [1] hull queue length [2] hull ticket head [3] degenerate [4] finished pair warps [5] pool fallbacks
  //        [8] pair CTAs that have started
  StepSync sy{ctx->d_ctr + 0, 0, 32, ctx->d_ctr + 8, ctx->d_ctr + 2, ctx->d_ctr + 4, ctx->d_ctr, ctx->d_prof, make_batch_plan(p.n_local, p.K, p.first_slots, 0, 0, 32), nullptr};
  auto set_plan = [&sy](const BatchPlan& b) { sy.plan = b; sy.narrow_pairs = b.narrow_a; sy.batch_width = b.width; };
  int pair_launches = 0, pair_grid = ctx->pair_blocks, hull_polling = 0;
  bool nbr_streamed = false;
  CU(cudaEventRecord(ctx->ev[0], st));
#define LQRO_LAUNCH_PAIRS(F, HT, PT, S, WARPS, PP, TB, ORD)                                                                   \
  k_pairs<F, HT, PT, S><<<pair_grid, (WARPS) * 32, ctx->pair_smem, pst>>>(PP, TB, d_x, d_nbr, po, sy, hq, ovf_base,             \
                                                                        ctx->pair_ovf_stride, ctx->pair_warps, ctx->d_hull_ws, ctx->hull_ws_stride, ORD)
#define LQRO_LAUNCH_PAIRS_ANY(PP, TB, ORD)                                                                                     \
  do {                                                                                                                         \
    if (ctx->fused_hull && !(ORD)) LQRO_LAUNCH_PAIRS(true, 0, 0, false, ctx->pair_warps + 1, PP, TB, ORD);                      \
    else if (p.H == 20 && p.P == 50) LQRO_LAUNCH_PAIRS(false, 20, 50, false, ctx->pair_warps, PP, TB, ORD); /* synthetic swarms */ \
    else if (p.H == 30 && p.P == 50) LQRO_LAUNCH_PAIRS(false, 30, 50, false, ctx->pair_warps, PP, TB, ORD); /* MAIN:30-34 */    \
    else LQRO_LAUNCH_PAIRS(false, 0, 0, false, ctx->pair_warps, PP, TB, ORD);                                                  \
  } while (0)
  /* the launch next to which the neighbour table arrives (sy.nbr_mark set, PP.seg_split > 0): the instantiations with the waits */
#define LQRO_LAUNCH_PAIRS_STREAMED(PP, TB)                                                                                     \
  do {                                                                                                                         \
    if (p.H == 20 && p.P == 50) LQRO_LAUNCH_PAIRS(false, 20, 50, true, ctx->pair_warps, PP, TB, (const int*)nullptr);           \
    else if (p.H == 30 && p.P == 50) LQRO_LAUNCH_PAIRS(false, 30, 50, true, ctx->pair_warps, PP, TB, (const int*)nullptr);      \
    else LQRO_LAUNCH_PAIRS(false, 0, 0, true, ctx->pair_warps, PP, TB, (const int*)nullptr);                                   \
  } while (0)
  cudaStream_t pst = st;  // stream of the pair-kernel launches
  double* ovf_base = ctx->d_pair_ovf;  // overflow clouds of the launch (one region per concurrent launch)
  if (ctx->h_nbr_chunked) {
    const int* h_nbr = ctx->h_nbr_chunked;
    ctx->h_nbr_chunked = nullptr;
    if (!ctx->copy_stream) {
      CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
      for (int c = 0; c < LQRO_NBR_CHUNKS; ++c) {
        CU(cudaStreamCreateWithFlags(&ctx->chunk_stream[c], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_up[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_done[c], cudaEventDisableTiming));
      }
      std::vector<int> id(p.n_local);
      for (int i = 0; i < p.n_local; ++i) id[i] = i;
      CU(cudaMalloc(&ctx->d_identity, sizeof(int) * (size_t)p.n_local));
      CU(cudaMemcpy(ctx->d_identity, id.data(), sizeof(int) * (size_t)p.n_local, cudaMemcpyHostToDevice));
      CU(cudaMalloc(&ctx->d_chunk_ctr, sizeof(int) * LQRO_NBR_CHUNKS));
    }
    CU(cudaMemsetAsync(ctx->d_chunk_ctr, 0, sizeof(int) * LQRO_NBR_CHUNKS, st));
    CU(cudaEventRecord(ctx->ev_ready, st));  // state gathered, counters cleared, previous step finished
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready, 0));  // d_nbr may still be read by the previous step
    const int per = (p.n_local + LQRO_NBR_CHUNKS - 1) / LQRO_NBR_CHUNKS;
    for (int c = 0; c < LQRO_NBR_CHUNKS; ++c) {
      const int off = c * per, cnt = std::min(per, p.n_local - off);
      if (cnt <= 0) break;
      CU(cudaMemcpyAsync(ctx->d_nbr + (size_t)off * p.K, h_nbr + (size_t)off * p.K, sizeof(int) * (size_t)cnt * p.K, cudaMemcpyHostToDevice, ctx->copy_stream));
      CU(cudaEventRecord(ctx->ev_up[c], ctx->copy_stream));
    }
    for (int c = 0; c < LQRO_NBR_CHUNKS; ++c) {
      const int off = c * per, cnt = std::min(per, p.n_local - off);
      if (cnt <= 0) break;
      pst = ctx->chunk_stream[c];
      CU(cudaStreamWaitEvent(pst, ctx->ev_up[c], 0));
      LqroParams pc = p;
      pc.n_local = cnt;
      set_plan(make_batch_plan(cnt, p.K, p.first_slots, 0, 0, 32));
      sy.batch_counter = ctx->d_chunk_ctr + c;
      const int* ord = ctx->d_identity + off;
      ovf_base = ctx->d_pair_ovf + (size_t)c * ctx->pair_ovf_stride * (size_t)ctx->pair_blocks * ctx->pair_warps;
      LQRO_LAUNCH_PAIRS_ANY(pc, ctx->tb, ord);
      CU(cudaEventRecord(ctx->ev_done[c], pst));
      CU(cudaStreamWaitEvent(st, ctx->ev_done[c], 0));
      ++pair_launches;
    }
    pst = st;
    ovf_base = ctx->d_pair_ovf;
    set_plan(make_batch_plan(p.n_local, p.K, p.first_slots, 0, 0, 32));
  } else if (!ctx->hetero && !ctx->fused_hull && ctx->overlap_sms > 0 && (size_t)p.n_local * p.K >= ((size_t)1 << 17)) {
    // LARGE SWARM: the hull queue gets a CONCURRENT CONSUMER.  A polling hull launch (k_hull, poll = 1) goes first, on its own
    // stream and on overlap_sms SMs that its shared-memory request fills; the pair kernel runs next to it on the other SMs and
    // hands out the nearest neighbour slots first (slot-major enumeration, narrow batches), so an inside pair reaches a hull
    // warp within the first tenth of the pair kernel and its hull -- a ~1 ms sequential job (qhull's insertion order) -- runs
    // under the rest of the pair kernel instead of behind it.  The wide launch after the pair kernel takes what is left.
    if (!ctx->hull_stream) {
      CU(cudaStreamCreateWithFlags(&ctx->hull_stream, cudaStreamNonBlocking));
      CU(cudaEventCreate(&ctx->ev_poll_go));
      CU(cudaEventCreate(&ctx->ev_poll1));
      CU(cudaEventCreate(&ctx->ev_poll0));
    }
    pair_grid = ctx->pair_blocks - ctx->overlap_sms;
    LqroParams pp = p;
    if (ctx->h_nbr_streamed) {
      // THE NEIGHBOUR TABLE COMES UP NEXT TO THE PAIR KERNEL (host-buffer steps, upload_step_inputs): its rows in two blocks on
      // their own stream -- the first twelfth of the agents, then the rest -- each followed by a 4-byte copy of "rows present" into
      // d_nbr_mark; the pair kernel enumerates the first block's agents first (LqroParams::seg_split) and waits for the mark
      // before it reads the rows of a batch (StepSync::nbr_mark).  The copies depend on nothing the kernel produces (they wait
      // for the state upload and the previous step only), so the wait always ends; the preferred velocities follow the table
      // and the LP waits for them.  65,536 x 64: 16.8 MB = 0.3 ms of PCIe time under the pair kernel instead of in front of it.
      const int* h_nbr = ctx->h_nbr_streamed;
      const double* h_vp = ctx->h_vpref_streamed;
      ctx->h_nbr_streamed = nullptr; ctx->h_vpref_streamed = nullptr;
      if (!ctx->nbr_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->nbr_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_nbr_go, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_nbr_up, cudaEventDisableTiming));
        CU(cudaMalloc(&ctx->d_nbr_mark, sizeof(int)));
        CU(cudaHostAlloc(&ctx->h_nbr_marks, 2 * sizeof(int), cudaHostAllocDefault));
      }
      // the first segment's pairs take 14 times longer to work on than the rest of the table takes to arrive (1.1 ns against
      // 0.08 ns per pair): a twelfth of the agents leaves a margin and delays the inside pairs of the others as little as possible
      const int n_a = (p.n_local / 12) & ~31;  // whole batches; n_local >= 2048 here
      ctx->h_nbr_marks[0] = n_a; ctx->h_nbr_marks[1] = p.n_local;
      CU(cudaMemsetAsync(ctx->d_nbr_mark, 0, sizeof(int), st));
      CU(cudaEventRecord(ctx->ev_nbr_go, st));  // state on the device, mark cleared, previous step finished (it read d_nbr)
      CU(cudaStreamWaitEvent(ctx->nbr_stream, ctx->ev_nbr_go, 0));
      const size_t row = sizeof(int) * (size_t)p.K;
      CU(cudaMemcpyAsync(ctx->d_nbr, h_nbr, row * n_a, cudaMemcpyHostToDevice, ctx->nbr_stream));
      CU(cudaMemcpyAsync(ctx->d_nbr_mark, ctx->h_nbr_marks + 0, sizeof(int), cudaMemcpyHostToDevice, ctx->nbr_stream));
      CU(cudaMemcpyAsync(ctx->d_nbr + (size_t)n_a * p.K, h_nbr + (size_t)n_a * p.K, row * (size_t)(p.n_local - n_a), cudaMemcpyHostToDevice, ctx->nbr_stream));
      CU(cudaMemcpyAsync(ctx->d_nbr_mark, ctx->h_nbr_marks + 1, sizeof(int), cudaMemcpyHostToDevice, ctx->nbr_stream));
      if (h_vp) CU(cudaMemcpyAsync(ctx->d_vpref, h_vp, sizeof(double) * 3 * (size_t)p.n_local, cudaMemcpyHostToDevice, ctx->nbr_stream));
      CU(cudaEventRecord(ctx->ev_nbr_up, ctx->nbr_stream));
      pp.seg_split = n_a;
      sy.nbr_mark = ctx->d_nbr_mark;
      nbr_streamed = true;
    }
    CU(cudaEventRecord(ctx->ev_poll_go, st));  // counters cleared, previous step finished
    CU(cudaStreamWaitEvent(ctx->hull_stream, ctx->ev_poll_go, 0));
    CU(cudaEventRecord(ctx->ev_poll0, ctx->hull_stream));
    k_hull<<<ctx->overlap_sms, 64 * ctx->hull_poll_warps, ctx->smem_optin, ctx->hull_stream>>>(
        p, ctx->tb, d_x, d_nbr, po, hq, ctx->d_ctr + 2, 1, ctx->d_ctr + 4, ctx->pair_warps * pair_grid, ctx->d_ctr + 8,
        ctx->d_hull_ws + ctx->hull_ws_stride * (size_t)ctx->hull_blocks, ctx->hull_ws_stride, ctx->d_ctr, ctx->d_prof, nullptr, ctx->d_tbs);
    CU(cudaEventRecord(ctx->ev_poll1, ctx->hull_stream));
    // batches of the nearest 8 slots (of either segment): narrow enough for >= 6 batches per pair warp of the grid
    set_plan(make_batch_plan(p.n_local, p.K, p.first_slots, pp.seg_split, 8, polling_batch_width(ctx)));
    if (nbr_streamed) LQRO_LAUNCH_PAIRS_STREAMED(pp, ctx->tb);
    else LQRO_LAUNCH_PAIRS_ANY(pp, ctx->tb, (const int*)nullptr);
    if (nbr_streamed) CU(cudaStreamWaitEvent(st, ctx->ev_nbr_up, 0));  // nothing to wait for once the pair kernel has ended, but the LP reads d_vpref
    pair_grid = ctx->pair_blocks;
    set_plan(make_batch_plan(p.n_local, p.K, p.first_slots, 0, 0, 32)); sy.nbr_mark = nullptr;
    ++pair_launches;
    hull_polling = 1;
  } else if (!ctx->hetero) {
    LQRO_LAUNCH_PAIRS_ANY(p, ctx->tb, (const int*)nullptr);
    ++pair_launches;
  } else {
    // one launch per dynamics class over the owned agents of that class (the tables of one class fill the CTA's shared memory)
    CU(cudaMemsetAsync(ctx->d_cls_ctr, 0, sizeof(int) * LQRO_MAX_CLASSES, st));
    for (int c = 0; c < LQRO_MAX_CLASSES; ++c) {
      if (!ctx->cls_count[c]) continue;
      LqroParams pc = p;
      pc.n_local = ctx->cls_count[c];
      set_plan(make_batch_plan(pc.n_local, p.K, p.first_slots, 0, 0, 32));
      LqroTables tbc = ctx->tb;
      tbc.NFt = ctx->cls[c].NFt; tbc.T = ctx->cls[c].T; tbc.CG = ctx->cls[c].CG; tbc.prn = ctx->cls[c].prn;
      sy.batch_counter = ctx->d_cls_ctr + c;
      const int* ord = ctx->d_order + ctx->cls_begin[c];
      LQRO_LAUNCH_PAIRS_ANY(pc, tbc, ord);
      ++pair_launches;
    }
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev[1], st));
  if (!hull_polling) k_hull_order<<<1, 1024, 0, st>>>(hq, ctx->d_ctr + 2, ctx->d_nreach, ctx->hull_blocks);  // more hulls than hull warps: longest first
  // the wide launch follows the pair kernel at once: the pollers stop taking tickets when the producers are done and finish the
  // hulls they hold, on their own global slots, next to it; the LP waits for both
  k_hull<<<ctx->hull_blocks, 64, ctx->hull_smem, st>>>(p, ctx->tb, d_x, d_nbr, po, hq, ctx->d_ctr + 2, 0, nullptr, 0, nullptr, ctx->d_hull_ws, ctx->hull_ws_stride,
                                                       ctx->d_ctr, ctx->d_prof, ctx->hetero ? ctx->d_agent_class : nullptr, ctx->d_tbs);
  CU(cudaGetLastError());
  if (hull_polling) CU(cudaStreamWaitEvent(st, ctx->ev_poll1, 0));
  CU(cudaEventRecord(ctx->ev[2], st));
  // one warp per agent whenever the planes of four agents fit the default shared memory: coalesced status / plane reads into
  // shared memory, the inner loops of the programs spread over the lanes (dense scenarios: O(K^2)); measured 0.047 ms against
  // 0.145 ms of the thread-per-agent kernel at 65,536 agents x 64 slots (LQRO_LP_WARP_MAX=<agents> restores the old split)
  static const int lp_warp_max = getenv("LQRO_LP_WARP_MAX") ? atoi(getenv("LQRO_LP_WARP_MAX")) : (1 << 30);
  if (p.n_local <= lp_warp_max && (size_t)4 * 12 * p.K * sizeof(float) <= 48 * 1024) {
    k_new_velocity_warp<<<(p.n_local + 3) / 4, 128, (size_t)4 * 12 * p.K * sizeof(float), st>>>(p.n_local, p.K, ctx->d_planes, ctx->d_status, d_vpref,
                                                                                                 p.lp_radius, d_newV, ctx->d_nplanes);
  } else {
    k_new_velocity<<<(p.n_local + 127) / 128, 128, 0, st>>>(p.n_local, p.K, ctx->d_planes, ctx->d_status, nullptr, d_vpref, p.lp_radius,
                                                            ctx->d_planes_c, ctx->d_proj, d_newV, ctx->d_nplanes);
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev[3], st));
  if (ctx->p2p && ctx->p2p_done_epoch < ctx->p2p_epoch && st == ctx->stream) {  // the step's last reader of the state is behind us: the peers may overwrite their slots
    k_p2p_done<<<1, 32, 0, st>>>(ctx->peers, ctx->rank, ctx->world, ctx->p2p_epoch);
    ctx->p2p_done_epoch = ctx->p2p_epoch;
  }
  ctx->last_polling = hull_polling != 0;
  ctx->last_nbr_streamed = nbr_streamed;
  ctx->launches = 2 + pair_launches + 1;  // pair kernel(s) + wide hull launch + LP + (polling hull launch | queue ordering)
  ctx->stepped = true;
  return LQRO_OK;
}

extern "C" int lqro_step_device(lqro_ctx* ctx, const double* d_x, const double* d_vpref, const int* d_nbr, double* d_newV, void* stream) {
  if (!ctx || !d_vpref || !d_newV) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_dynamics) return fail(ctx, LQRO_E_STATE, "set_dynamics first");
  if (!d_nbr && !ctx->have_nbr) return fail(ctx, LQRO_E_STATE, "no neighbour table: pass d_nbr, or step once with a host table / lqro_knn_device first");
  CU(cudaSetDevice(ctx->cfg.device));
  ctx->h_nbr_chunked = nullptr; ctx->h_nbr_streamed = nullptr; ctx->h_vpref_streamed = nullptr;  // device-resident inputs: nothing comes from the host
  return launch_step(ctx, d_x ? d_x : ctx->d_x, d_vpref, d_nbr ? d_nbr : ctx->d_nbr, d_newV, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int lqro_step(lqro_ctx* ctx, const double* x, const double* vpref, const int* nbr, double* newV) {
  if (!ctx || !x || !vpref || !newV) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_dynamics) return fail(ctx, LQRO_E_STATE, "set_dynamics first");
  if (!nbr && !ctx->device_knn && !ctx->have_nbr) return fail(ctx, LQRO_E_STATE, "no neighbour table: the first step needs nbr (or LQRO_FLAG_DEVICE_KNN)");
  if (nbr) ctx->have_nbr = true;
  CU(cudaSetDevice(ctx->cfg.device));
  const LqroParams& p = ctx->prm;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(ctx->d_x, x, sizeof(double) * 16 * (size_t)p.N, cudaMemcpyHostToDevice, st));
  // large swarms: the neighbour table goes up next to the pair kernel (see upload_step_inputs, launch_step)
  int rc = upload_step_inputs(ctx, vpref, nbr, st);
  if (rc) return rc;
  if (!nbr && ctx->device_knn && (rc = launch_knn(ctx, ctx->d_x, ctx->d_nbr, st))) return rc;
  rc = launch_step(ctx, ctx->d_x, ctx->d_vpref, ctx->d_nbr, ctx->d_newV, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(newV, ctx->d_newV, sizeof(double) * 3 * (size_t)p.n_local, cudaMemcpyDeviceToHost, st));
  int ctr9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // d_ctr[1..9]: [0] hull queue length ... [8] neighbour rows that never arrived
  CU(cudaMemcpyAsync(ctr9, ctx->d_ctr + 1, sizeof(ctr9), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const int queued = ctr9[0];
  if (ctr9[8]) return fail(ctx, LQRO_E_CUDA, "the neighbour table did not arrive on the device while the pair kernel waited for it");
  if (queued > ctx->hull_cap) return fail(ctx, LQRO_E_LIMIT, "hull queue overflow: inside pairs beyond max_hull_pairs got no half-plane (status 0x200)");
  return LQRO_OK;
}

extern "C" int lqro_set_stream(lqro_ctx* ctx, void* stream) {
  if (!ctx) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
  return LQRO_OK;
}

extern "C" int lqro_synchronize(lqro_ctx* ctx) {
  if (!ctx) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_get_pair_results(lqro_ctx* ctx, float* planes, int* status, int* nreach, double* dist, int* nfacets) {
  if (!ctx) return LQRO_E_ARG;
  if (!ctx->stepped) return fail(ctx, LQRO_E_STATE, "no step yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  const size_t NP = (size_t)ctx->prm.n_local * ctx->prm.K;
  if (planes) CU(cudaMemcpy(planes, ctx->d_planes, sizeof(float) * 6 * NP, cudaMemcpyDeviceToHost));
  if (status) CU(cudaMemcpy(status, ctx->d_status, sizeof(int) * NP, cudaMemcpyDeviceToHost));
  if (nreach) CU(cudaMemcpy(nreach, ctx->d_nreach, sizeof(int) * NP, cudaMemcpyDeviceToHost));
  if (dist) CU(cudaMemcpy(dist, ctx->d_dist, sizeof(double) * NP, cudaMemcpyDeviceToHost));
  if (nfacets) CU(cudaMemcpy(nfacets, ctx->d_nfacets, sizeof(int) * NP, cudaMemcpyDeviceToHost));
  return LQRO_OK;
}

extern "C" int lqro_get_counters(lqro_ctx* ctx, long long* out6) {
  if (!ctx || !out6) return LQRO_E_ARG;
  if (!ctx->stepped) return fail(ctx, LQRO_E_STATE, "no step yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  int c[8];
  CU(cudaMemcpy(c, ctx->d_ctr, sizeof c, cudaMemcpyDeviceToHost));
  const size_t NP = (size_t)ctx->prm.n_local * ctx->prm.K;
  std::vector<int> st(NP);
  CU(cudaMemcpy(st.data(), ctx->d_status, sizeof(int) * NP, cudaMemcpyDeviceToHost));
  long long withplane = 0;
  for (size_t i = 0; i < NP; ++i) withplane += (st[i] & 0xff) != 0;
  out6[0] = (long long)NP;
  out6[1] = withplane;
  out6[2] = c[1];
  out6[3] = c[3];
  out6[4] = c[1] > ctx->hull_cap ? c[1] - ctx->hull_cap : 0;
  out6[5] = ctx->launches;
  return LQRO_OK;
}

// ---- simulator-form agent update (lqro_sim.cuh) -----------------------------------------------
static_assert(sizeof(lqro_sim_model) == sizeof(LqroSimModel), "lqro_sim_model and LqroSimModel must have one layout");

extern "C" void lqro_sim_default_model(lqro_sim_model* m) {  // setup(), MAIN:176-198; M, N MAIN:1139-1140
  memset(m, 0, sizeof *m);
  m->dt = 1.0 / 30.0;
  m->gravity = 9.80665;
  m->mass = 0.500;
  m->moment_const = 1.5e-9 / 6.11e-8;
  m->thrust_latency = 40.0;
  m->length = 0.3429 / 2;
  m->jstep = 0.0009765625;
  // inertia = 0.1*identity<3>(); invInertia = !inertia (full-pivot Gauss of a diagonal matrix: 1/0.1 on the diagonal)
  for (int i = 0; i < 3; ++i) { m->inertia[4 * i] = 1.0 * 0.1; m->inv_inertia[4 * i] = 1.0 / (1.0 * 0.1); }
  const double nominal = m->gravity * m->mass / 4;
  for (int i = 0; i < 4; ++i) m->u_goal[i] = nominal;
  for (int i = 0; i < 16; ++i) m->M[17 * i] = 1.0 * 0.000000001;
  for (int i = 0; i < 6; ++i) m->N[7 * i] = 1.0 * 0.000000001;
}

extern "C" int lqro_sim_set_model(lqro_ctx* ctx, const lqro_sim_model* m) {
  if (!ctx || !m) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  if (!ctx->d_sim_model) CU(cudaMalloc(&ctx->d_sim_model, sizeof(LqroSimModel)));
  CU(cudaMemcpyAsync(ctx->d_sim_model, m, sizeof(LqroSimModel), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->sim_m_diag = 1;  // diagonal motion-noise covariance: two of the four products of the noise covariance collapse (lqro_sim.cuh)
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) if (i != j && m->M[i * 16 + j] != 0.0) ctx->sim_m_diag = 0;
  ctx->have_sim_model = true;
  return LQRO_OK;
}

// Four launches per update (lqro_sim.cuh): the control input (thread per agent), the two linearised flows of every agent as
// independent warp tasks (11 KB of shared memory per warp, 20 warps per SM), the process-noise sample (four agents per warp, 13 KB per warp,
// 16 warps per SM), the state / covariance updates + observation + kalmanFilter2 + findVGoal (11.3 KB, 20 warps per SM).
static int launch_sim_update(lqro_ctx* ctx, int n, double* x, double* Rot, double* P, double* xTrue, double* RotTrue, const double* newV,
                             const double* pGoal, const double* normals, double* vGoal, double* u, cudaStream_t st) {
  const int wa = SIM_FLOW_WARPS, wb = 4, wc = 4;
  const size_t smem_a = (size_t)wa * sizeof(LqroSimWsA), smem_b = (size_t)wb * 4 * SIM_NOISE_GROUP_DOUBLES * sizeof(double), smem_c = (size_t)wc * sizeof(LqroSimWsC);
  if (!ctx->sim_attr_set) {  // per device, hence per context
    CU(cudaFuncSetAttribute(k_sim_flow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    CU(cudaFuncSetAttribute(k_sim_noise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    CU(cudaFuncSetAttribute(k_sim_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    ctx->sim_attr_set = true;
  }
  for (cudaEvent_t& e : ctx->ev_sim) if (!e) CU(cudaEventCreate(&e));
  const size_t per = 4 + 2 * 256 + 2 * 16 + 256 + 16;  // u | MM x2 | v1 x2 | Ae | samp   (doubles per agent)
  const size_t need = sizeof(double) * per * (size_t)n;
  if (ctx->sim_scratch_bytes < need) {
    CU(cudaStreamSynchronize(st));
    if (ctx->d_sim_scratch) cudaFree(ctx->d_sim_scratch);
    ctx->d_sim_scratch = nullptr; ctx->sim_scratch_bytes = 0;
    CU(cudaMalloc(&ctx->d_sim_scratch, need));
    ctx->sim_scratch_bytes = need;
  }
  const size_t N = (size_t)n;
  LqroSimScratch sc;
  sc.MM = (double*)ctx->d_sim_scratch; sc.Ae = sc.MM + 512 * N; sc.v1 = sc.Ae + 256 * N; sc.samp = sc.v1 + 32 * N; sc.u = sc.samp + 16 * N;
  auto grid = [&](long long items, int warps, int ctas_per_sm) {
    long long b = (items + warps - 1) / warps; const long long cap = (long long)ctx->sm_count * ctas_per_sm; return (int)(b > cap ? cap : b); };
  CU(cudaEventRecord(ctx->ev_sim[0], st));
  k_sim_control<<<(n + 127) / 128, 128, 0, st>>>(n, ctx->d_sim_model, x, Rot, newV, sc.u, u);
  CU(cudaGetLastError());
  k_sim_flow<<<grid(2LL * n, wa, SIM_FLOW_CTAS), wa * 32, smem_a, st>>>(n, ctx->d_sim_model, x, Rot, xTrue, RotTrue, sc, ctx->sim_m_diag);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev_sim[1], st));
  k_sim_noise<<<grid((n + 3) / 4, wb, 4), wb * 32, smem_b, st>>>(n, normals, sc);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev_sim[2], st));
  k_sim_filter<<<grid(n, wc, SIM_FILTER_CTAS), wc * 32, smem_c, st>>>(n, ctx->d_sim_model, x, Rot, P, xTrue, RotTrue, pGoal, normals, sc, vGoal);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev_sim[3], st));
  ctx->sim_launches += 4;
  return LQRO_OK;
}

extern "C" int lqro_sim_update_device(lqro_ctx* ctx, int n, double* d_x, double* d_Rot, double* d_P, double* d_xTrue, double* d_RotTrue,
                                      const double* d_newV, const double* d_pGoal, const double* d_normals, double* d_vGoal, double* d_u,
                                      void* stream) {
  if (!ctx || n < 0 || !d_x || !d_Rot || !d_P || !d_xTrue || !d_RotTrue || !d_newV || !d_pGoal || !d_normals || !d_vGoal)
    return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_sim_model) return fail(ctx, LQRO_E_STATE, "lqro_sim_set_model first");
  if (n == 0) return LQRO_OK;
  CU(cudaSetDevice(ctx->cfg.device));
  return launch_sim_update(ctx, n, d_x, d_Rot, d_P, d_xTrue, d_RotTrue, d_newV, d_pGoal, d_normals, d_vGoal, d_u,
                           stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int lqro_sim_update(lqro_ctx* ctx, int n, double* x, double* Rot, double* P, double* xTrue, double* RotTrue, const double* newV,
                               const double* pGoal, const double* normals, double* vGoal, double* u) {
  if (!ctx || n < 0 || !x || !Rot || !P || !xTrue || !RotTrue || !newV || !pGoal || !normals || !vGoal) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_sim_model) return fail(ctx, LQRO_E_STATE, "lqro_sim_set_model first");
  if (n == 0) return LQRO_OK;
  CU(cudaSetDevice(ctx->cfg.device));
  // one device block: x 16 | Rot 9 | P 256 | xTrue 16 | RotTrue 9 | newV 3 | pGoal 3 | normals 22 | vGoal 3 | u 4  (doubles per agent)
  const size_t per = 16 + 9 + 256 + 16 + 9 + 3 + 3 + 22 + 3 + 4;
  const size_t need = sizeof(double) * per * (size_t)n;
  if (ctx->sim_buf_bytes < need) {
    if (ctx->d_sim_buf) cudaFree(ctx->d_sim_buf);
    ctx->d_sim_buf = nullptr; ctx->sim_buf_bytes = 0;
    CU(cudaMalloc(&ctx->d_sim_buf, need));
    ctx->sim_buf_bytes = need;
  }
  double* b = (double*)ctx->d_sim_buf;
  const size_t N = (size_t)n;
  double *dx = b, *dR = dx + 16 * N, *dP = dR + 9 * N, *dxT = dP + 256 * N, *dRT = dxT + 16 * N, *dnV = dRT + 9 * N, *dpG = dnV + 3 * N,
         *dnr = dpG + 3 * N, *dvG = dnr + 22 * N, *du = dvG + 3 * N;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(dx, x, 8 * 16 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dR, Rot, 8 * 9 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dP, P, 8 * 256 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dxT, xTrue, 8 * 16 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dRT, RotTrue, 8 * 9 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dnV, newV, 8 * 3 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dpG, pGoal, 8 * 3 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dnr, normals, 8 * 22 * N, cudaMemcpyHostToDevice, st));
  int rc = launch_sim_update(ctx, n, dx, dR, dP, dxT, dRT, dnV, dpG, dnr, dvG, du, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(x, dx, 8 * 16 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(Rot, dR, 8 * 9 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(P, dP, 8 * 256 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(xTrue, dxT, 8 * 16 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(RotTrue, dRT, 8 * 9 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(vGoal, dvG, 8 * 3 * N, cudaMemcpyDeviceToHost, st));
  if (u) CU(cudaMemcpyAsync(u, du, 8 * 4 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LQRO_OK;
}

extern "C" int lqro_sim_last_ms(lqro_ctx* ctx, float* out4) {
  if (!ctx || !out4) return LQRO_E_ARG;
  if (!ctx->ev_sim[0]) return fail(ctx, LQRO_E_STATE, "no simulator update yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaEventSynchronize(ctx->ev_sim[3]));
  CU(cudaEventElapsedTime(out4 + 0, ctx->ev_sim[0], ctx->ev_sim[3]));
  for (int i = 0; i < 3; ++i) CU(cudaEventElapsedTime(out4 + 1 + i, ctx->ev_sim[i], ctx->ev_sim[i + 1]));
  return LQRO_OK;
}

extern "C" int lqro_debug_sim_cycles(lqro_ctx* ctx, long long* out16, int reset) {
  if (!ctx || !out16) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpyFromSymbol(out16, g_sim_prof, sizeof(long long) * 16));
  if (reset) { long long z[16] = {0}; CU(cudaMemcpyToSymbol(g_sim_prof, z, sizeof z)); }
  return LQRO_OK;
}

// Host only: the reference's normal() (MAIN:325-336) over the MSVC CRT generator it was written for.
extern "C" int lqro_msvc_normals(unsigned int* seed_io, int count, double* out) {
  if (!seed_io || count < 0 || (count && !out)) return LQRO_E_ARG;
  unsigned int seed = *seed_io;
  auto rnd = [&seed]() -> int { seed = seed * 214013u + 2531011u; return (int)((seed >> 16) & 0x7fff); };
  const int RM = 32767;
  auto uniform = [&]() -> double {  // random(), MAIN:320-322: integer arithmetic, then one division
    const int a = rnd();
    const int b = rnd();
    return (double)(a * (RM + 1) + b) / (RM * (RM + 2));
  };
  for (int i = 0; i < count; ++i) {
    double u = 0, v = 0, s = 0;
    while (s == 0 || s > 1) {
      u = 2 * uniform() - 1;
      v = 2 * uniform() - 1;
      s = u * u + v * v;
    }
    out[i] = u * sqrt(-2 * log(s) / s);
  }
  *seed_io = seed;
  return LQRO_OK;
}

extern "C" int lqro_debug_hull_cycles(lqro_ctx* ctx, long long* out16) {
  if (!ctx || !out16) return LQRO_E_ARG;
  if (!ctx->stepped) return fail(ctx, LQRO_E_STATE, "no step yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
#ifdef LQRO_PAIR_PROF  // profiling build: the pair kernel's phase cycles instead
  CU(cudaMemcpy(out16, ctx->d_prof + 16, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
#else
  CU(cudaMemcpy(out16, ctx->d_prof, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
#endif
  return LQRO_OK;
}

extern "C" int lqro_debug_hull_longest(lqro_ctx* ctx, long long* out16) {  // profiling builds: phase cycles of the longest hull of the last step, [12] points, [13] facets
  if (!ctx || !out16) return LQRO_E_ARG;
  if (!ctx->stepped) return fail(ctx, LQRO_E_STATE, "no step yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(out16, ctx->d_prof + 16, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
  return LQRO_OK;
}

extern "C" int lqro_last_step_detail_ms(lqro_ctx* ctx, float* out5) {
  float* out4 = out5;
  if (!ctx || !out5) return LQRO_E_ARG;
  if (!ctx->stepped) return fail(ctx, LQRO_E_STATE, "no step yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaEventSynchronize(ctx->ev[3]));
  out4[0] = out4[1] = out4[2] = 0.f; out4[3] = (float)ctx->overlap_sms; out5[4] = (float)ctx->prm.first_slots;
  if (!ctx->last_polling) { CU(cudaEventElapsedTime(out4 + 1, ctx->ev[0], ctx->ev[1])); out4[3] = 0.f; return LQRO_OK; }
  CU(cudaEventElapsedTime(out4 + 0, ctx->ev[0], ctx->ev_poll0));  // until the polling launch is issued on its stream
  CU(cudaEventElapsedTime(out4 + 1, ctx->ev_poll_go, ctx->ev[1]));
  CU(cudaEventElapsedTime(out4 + 2, ctx->ev_poll0, ctx->ev_poll1));
  return LQRO_OK;
}

extern "C" int lqro_last_step_nbr_streamed(lqro_ctx* ctx) { return ctx && ctx->stepped && ctx->last_nbr_streamed ? 1 : 0; }

extern "C" int lqro_last_step_ms(lqro_ctx* ctx, float* out4) {
  if (!ctx || !out4) return LQRO_E_ARG;
  if (!ctx->stepped) return fail(ctx, LQRO_E_STATE, "no step yet");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaEventSynchronize(ctx->ev[3]));
  CU(cudaEventElapsedTime(out4 + 0, ctx->ev[0], ctx->ev[1]));
  CU(cudaEventElapsedTime(out4 + 1, ctx->ev[1], ctx->ev[2]));
  CU(cudaEventElapsedTime(out4 + 2, ctx->ev[2], ctx->ev[3]));
  CU(cudaEventElapsedTime(out4 + 3, ctx->ev[0], ctx->ev[3]));
  return LQRO_OK;
}

// ---- per-stage entry points ---------------------------------------------------------------------------
extern "C" int lqro_candidates(lqro_ctx* ctx, const double* xi, const double* xj, double* q) {
  if (!ctx || !xi || !xj || !q) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_dynamics) return fail(ctx, LQRO_E_STATE, "set_dynamics first");
  CU(cudaSetDevice(ctx->cfg.device));
  const int HP = ctx->prm.HP;
  int rc = ensure_scratch(ctx, sizeof(double) * (32 + 3 * (size_t)HP));
  if (rc) return rc;
  double* dxi = (double*)ctx->d_scratch; double* dxj = dxi + 16; double* dq = dxj + 16;
  CU(cudaMemcpyAsync(dxi, xi, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dxj, xj, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
  const size_t smem = sizeof(double) * (pair_block_table_doubles(ctx->prm.H, ctx->prm.P) + pair_warp_doubles(ctx->prm.H, ctx->prm.P));
  k_stage_candidates<<<1, 32, smem, ctx->stream>>>(ctx->prm, ctx->tb, dxi, dxj, dq);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(q, dq, sizeof(double) * 3 * HP, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_find_fg(lqro_ctx* ctx, const double* A, const double* B, const double* L, const double* E, double* F, double* G) {
  if (!ctx || !A || !B || !L || !E || !F || !G) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  int rc = ensure_scratch(ctx, sizeof(double) * (256 + 64 + 64 + 12 + 256 + 48));
  if (rc) return rc;
  double* d = (double*)ctx->d_scratch;
  double *dA = d, *dB = d + 256, *dL = dB + 64, *dE = dL + 64, *dF = dE + 12, *dG = dF + 256;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(dA, A, sizeof(double) * 256, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dB, B, sizeof(double) * 64, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dL, L, sizeof(double) * 64, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dE, E, sizeof(double) * 12, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dF, F, sizeof(double) * 256, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dG, G, sizeof(double) * 48, cudaMemcpyHostToDevice, st));
  k_stage_findfg<<<1, 256, 0, st>>>(dA, dB, dL, dE, dF, dG);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(F, dF, sizeof(double) * 256, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(G, dG, sizeof(double) * 48, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LQRO_OK;
}

extern "C" int lqro_create_obstacle(lqro_ctx* ctx, int n_points, const double* points, const double* F, const double* G, const double* P,
                                    const double* x1, const double* x2, double* Transform, double* Translate, double* q) {
  if (!ctx || n_points < 1 || !points || !F || !G || !P || !x1 || !x2 || !Transform || !Translate || !q) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  const size_t n3 = 3 * (size_t)n_points;
  int rc = ensure_scratch(ctx, sizeof(double) * (2 * n3 + 256 + 48 + 48 + 32 + 12));
  if (rc) return rc;
  double* d = (double*)ctx->d_scratch;
  double *dp = d, *dq = dp + n3, *dF = dq + n3, *dG = dF + 256, *dP = dG + 48, *dx1 = dP + 48, *dx2 = dx1 + 16, *dT = dx2 + 16, *dTr = dT + 9;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(dp, points, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dF, F, sizeof(double) * 256, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dG, G, sizeof(double) * 48, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dP, P, sizeof(double) * 48, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dx1, x1, sizeof(double) * 16, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dx2, x2, sizeof(double) * 16, cudaMemcpyHostToDevice, st));
  k_stage_create_obstacle<<<1, 32, 0, st>>>(n_points, dp, dF, dG, dP, dx1, dx2, dT, dTr, dq);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(q, dq, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(Transform, dT, sizeof(double) * 9, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(Translate, dTr, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LQRO_OK;
}

extern "C" int lqro_reachable(lqro_ctx* ctx, const double* q, int n, const double* xi, const double* xj, int* idx, int* n_out) {
  if (!ctx || !q || n < 0 || !xi || !xj || !idx || !n_out) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  int rc = ensure_scratch(ctx, sizeof(double) * (32 + 3 * (size_t)n) + sizeof(int) * ((size_t)n + 4));
  if (rc) return rc;
  double* dxi = (double*)ctx->d_scratch; double* dxj = dxi + 16; double* dq = dxj + 16;
  int* didx = (int*)(dq + 3 * (size_t)n); int* dn = didx + n;
  CU(cudaMemcpyAsync(dxi, xi, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dxj, xj, sizeof(double) * 16, cudaMemcpyHostToDevice, ctx->stream));
  if (n) CU(cudaMemcpyAsync(dq, q, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
  k_stage_reachable<<<1, 32, 0, ctx->stream>>>(ctx->prm, dq, n, dxi, dxj, didx, dn);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(n_out, dn, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (n) CU(cudaMemcpyAsync(idx, didx, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_gjk(lqro_ctx* ctx, const double* vrel, const double* pts, int n, double* sqdist, double* wpt_vrel, double* wpt_hull,
                        int* iterations) {
  if (!ctx || !vrel || !pts || n < 1 || !sqdist) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  int rc = ensure_scratch(ctx, sizeof(double) * (16 + 6 * (size_t)n) + 16);
  if (rc) return rc;
  double* dv = (double*)ctx->d_scratch; double* dout = dv + 4; double* dp = dout + 8; double* dscr = dp + 3 * (size_t)n;
  int* dit = (int*)(dscr + 3 * (size_t)n);
  CU(cudaMemcpyAsync(dv, vrel, sizeof(double) * 3, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dp, pts, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
  k_stage_gjk<<<1, 32, 0, ctx->stream>>>(dv, dp, n, dscr, dout, dit);
  CU(cudaGetLastError());
  double out[8]; int it = 0;
  CU(cudaMemcpyAsync(out, dout, sizeof(double) * 7, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&it, dit, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *sqdist = out[0];
  if (wpt_hull) { wpt_hull[0] = out[1]; wpt_hull[1] = out[2]; wpt_hull[2] = out[3]; }
  if (wpt_vrel) { wpt_vrel[0] = out[4]; wpt_vrel[1] = out[5]; wpt_vrel[2] = out[6]; }
  if (iterations) *iterations = it;
  return LQRO_OK;
}

static int run_stage_hull(lqro_ctx* ctx, const double* pts, int n, int mode, double* planes, int* verts, int maxf, int* info5,
                          const double* vrel, double* escape4) {
  if (n > ctx->stage_ws_pts) {
    if (ctx->d_stage_ws) cudaFree(ctx->d_stage_ws);
    ctx->d_stage_ws = nullptr; ctx->stage_ws_pts = 0;
    CU(cudaMalloc(&ctx->d_stage_ws, hull_ws_bytes(n)));
    ctx->stage_ws_pts = n;
  }
  const int cap = ctx->stage_ws_pts;
  const size_t nf_cap = 2 * (size_t)n + 64;
  int rc = ensure_scratch(ctx, sizeof(double) * (6 * (size_t)n + 4 * nf_cap + 16) + sizeof(int) * (3 * nf_cap + 16));
  if (rc) return rc;
  double* dp = (double*)ctx->d_scratch; double* dpl = dp + 6 * (size_t)n;
  double* dv = dpl + 4 * nf_cap; double* desc = dv + 4;
  int* dvt = (int*)(desc + 4); int* dinfo = dvt + 3 * nf_cap;
  CU(cudaMemcpyAsync(dp, pts, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
  if (vrel) CU(cudaMemcpyAsync(dv, vrel, sizeof(double) * 3, cudaMemcpyHostToDevice, ctx->stream));
  k_stage_hull<<<1, 32, 0, ctx->stream>>>(dp, n, mode, ctx->d_stage_ws, cap, dpl, dvt, (int)nf_cap, dinfo, vrel ? dv : nullptr, desc);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(info5, dinfo, sizeof(int) * 5, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  int nf = info5[0];
  if (nf > maxf) nf = maxf;
  if (planes && nf > 0) CU(cudaMemcpy(planes, dpl, sizeof(double) * 4 * nf, cudaMemcpyDeviceToHost));
  if (verts && nf > 0) CU(cudaMemcpy(verts, dvt, sizeof(int) * 3 * nf, cudaMemcpyDeviceToHost));
  if (escape4 && vrel) CU(cudaMemcpy(escape4, desc, sizeof(double) * 4, cudaMemcpyDeviceToHost));
  return LQRO_OK;
}

extern "C" int lqro_convex_hull(lqro_ctx* ctx, const double* pts, int n, double* planes, int* verts, int* nf_io, int* degenerate) {
  if (!ctx || !pts || n < 4 || !nf_io) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  int info[5];
  int rc = run_stage_hull(ctx, pts, n, 0, planes, verts, *nf_io, info, nullptr, nullptr);
  if (rc) return rc;
  if (degenerate) *degenerate = info[1];
  if (info[2]) { *nf_io = 0; return fail(ctx, LQRO_E_LIMIT, "hull failed (flat input or internal limit)"); }
  *nf_io = info[0];
  return LQRO_OK;
}

extern "C" int lqro_hull_escape(lqro_ctx* ctx, const double* pts, int n, const double* vrel, double* distance, double* normal,
                                int* best_facet, int* n_facets) {
  if (!ctx || !pts || n < 4 || !vrel || !distance || !normal) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  int info[5]; double esc[4];
  int rc = run_stage_hull(ctx, pts, n, ctx->prm.hull_input == LQRO_HULL_TEXT6 ? 1 : 0, nullptr, nullptr, 0, info, vrel, esc);
  if (rc) return rc;
  if (info[2]) return fail(ctx, LQRO_E_LIMIT, "hull failed (flat input or internal limit)");
  *distance = esc[0]; normal[0] = esc[1]; normal[1] = esc[2]; normal[2] = esc[3];
  if (best_facet) *best_facet = info[4];
  if (n_facets) *n_facets = info[0];
  return LQRO_OK;
}

extern "C" int lqro_new_velocity(lqro_ctx* ctx, int batch, int max_planes, const float* planes, const int* counts, const double* goal,
                                 double* newV) {
  if (!ctx || batch < 1 || max_planes < 1 || !planes || !counts || !goal || !newV) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  const size_t np = (size_t)batch * max_planes * 6;
  int rc = ensure_scratch(ctx, sizeof(float) * 3 * np + sizeof(double) * 6 * (size_t)batch + sizeof(int) * (size_t)batch + 64);
  if (rc) return rc;
  double* dgoal = (double*)ctx->d_scratch; double* dnew = dgoal + 3 * (size_t)batch;
  float* dpl = (float*)(dnew + 3 * (size_t)batch); float* dpc = dpl + np; float* dpr = dpc + np;
  int* dcnt = (int*)(dpr + np);
  CU(cudaMemcpyAsync(dpl, planes, sizeof(float) * np, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dcnt, counts, sizeof(int) * batch, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dgoal, goal, sizeof(double) * 3 * batch, cudaMemcpyHostToDevice, ctx->stream));
  k_new_velocity<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(batch, max_planes, dpl, nullptr, dcnt, dgoal, ctx->prm.lp_radius, dpc, dpr, dnew, nullptr);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(newV, dnew, sizeof(double) * 3 * batch, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_agent_update(lqro_ctx* ctx, int n, float* pos, float* vel, const float* newVel, float dt) {
  if (!ctx || n < 1 || !pos || !vel || !newVel) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  const size_t n3 = 3 * (size_t)n;
  int rc = ensure_scratch(ctx, sizeof(float) * 3 * n3);
  if (rc) return rc;
  float* dp = (float*)ctx->d_scratch; float* dv = dp + n3; float* dn = dv + n3;
  CU(cudaMemcpyAsync(dp, pos, sizeof(float) * n3, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dn, newVel, sizeof(float) * n3, cudaMemcpyHostToDevice, ctx->stream));
  k_agent_update<<<(int)((n3 + 255) / 256), 256, 0, ctx->stream>>>((int)n3, dp, dv, dn, dt);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(pos, dp, sizeof(float) * n3, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(vel, dv, sizeof(float) * n3, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_round6(lqro_ctx* ctx, const double* in, int n, double* out) {
  if (!ctx || !in || !out || n < 1) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  int rc = ensure_scratch(ctx, sizeof(double) * 2 * (size_t)n);
  if (rc) return rc;
  double* di = (double*)ctx->d_scratch; double* dout = di + n;
  CU(cudaMemcpyAsync(di, in, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  k_round6<<<(n + 255) / 256, 256, 0, ctx->stream>>>(di, n, dout);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LQRO_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------------------
extern "C" int lqro_comm_unique_id(void* id128) {
  if (!id128) return LQRO_E_ARG;
  if (!load_nccl()) { g_create_err = "libnccl not found"; return LQRO_E_NCCL; }
  lqro_ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return LQRO_E_NCCL;
  memcpy(id128, &id, 128);
  return LQRO_OK;
}

// Peer-memory all-gather (lqro_p2p.cuh): map every peer's state buffer and flag words into this process.  The CUDA IPC handles
// travel over the NCCL communicator that has just been built (one all-gather of 136 bytes per rank), then every rank reports whether
// it could open all of them and the path is used only if ALL ranks could -- otherwise (or with LQRO_NO_P2P_GATHER=1, or more than
// LQRO_P2P_MAX_RANKS ranks) the step keeps ncclAllGather.  Never an error: the fallback is always there.
static int p2p_setup(lqro_ctx* ctx) {
  const int world = ctx->world, rank = ctx->rank;
  ctx->p2p = false;
  for (int r = 0; r < LQRO_P2P_MAX_RANKS; ++r) { ctx->peers.x[r] = nullptr; ctx->peers.flags[r] = nullptr; }
  if (world < 2) return LQRO_OK;
  struct Rec { cudaIpcMemHandle_t hx, hf; int want; int pad; };
  static_assert(sizeof(Rec) % 8 == 0, "record size");
  const char* off = getenv("LQRO_NO_P2P_GATHER");
  int want = (world <= LQRO_P2P_MAX_RANKS && !(off && atoi(off) != 0)) ? 1 : 0;
  Rec mine;
  memset(&mine, 0, sizeof mine);
  if (want) {
    if (cudaMalloc(&ctx->d_p2p_flags, sizeof(int) * LQRO_P2P_FLAG_INTS) != cudaSuccess) want = 0;
    else {
      cudaMemset(ctx->d_p2p_flags, 0, sizeof(int) * LQRO_P2P_FLAG_INTS);
      cudaDeviceSynchronize();
      if (cudaIpcGetMemHandle(&mine.hx, ctx->d_x) != cudaSuccess || cudaIpcGetMemHandle(&mine.hf, ctx->d_p2p_flags) != cudaSuccess) want = 0;
    }
  }
  (void)cudaGetLastError();
  mine.want = want;
  // every rank takes part in the two exchanges whatever it wants itself: they are collectives
  char* d_rec = nullptr;
  CU(cudaMalloc(&d_rec, sizeof(Rec) * (size_t)world));
  CU(cudaMemcpy(d_rec + sizeof(Rec) * (size_t)rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
  const int ncclInt8 = 0;
  int rc = g_nccl.AllGather(d_rec + sizeof(Rec) * (size_t)rank, d_rec, sizeof(Rec), ncclInt8, ctx->comm, ctx->stream);
  if (rc != 0) { cudaFree(d_rec); return fail(ctx, LQRO_E_NCCL, "ncclAllGather of the IPC handles failed"); }
  std::vector<Rec> all(world);
  CU(cudaMemcpyAsync(all.data(), d_rec, sizeof(Rec) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  int ok = 1;
  for (int r = 0; r < world; ++r) ok = ok && all[r].want;
  if (ok) {
    ctx->peers.x[rank] = ctx->d_x; ctx->peers.flags[rank] = ctx->d_p2p_flags;
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      void *px = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&px, all[r].hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
      ctx->peers.x[r] = (double*)px;
      if (cudaIpcOpenMemHandle(&pf, all[r].hf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
      ctx->peers.flags[r] = (int*)pf;
    }
    (void)cudaGetLastError();
  }
  // second exchange: could everybody open everything?
  Rec vote;
  memset(&vote, 0, sizeof vote);
  vote.want = ok;
  CU(cudaMemcpy(d_rec + sizeof(Rec) * (size_t)rank, &vote, sizeof(Rec), cudaMemcpyHostToDevice));
  rc = g_nccl.AllGather(d_rec + sizeof(Rec) * (size_t)rank, d_rec, sizeof(Rec), ncclInt8, ctx->comm, ctx->stream);
  if (rc != 0) { cudaFree(d_rec); return fail(ctx, LQRO_E_NCCL, "ncclAllGather of the IPC votes failed"); }
  CU(cudaMemcpyAsync(all.data(), d_rec, sizeof(Rec) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_rec);
  int all_ok = 1;
  for (int r = 0; r < world; ++r) all_ok = all_ok && all[r].want;
  if (!all_ok) {
    for (int r = 0; r < world; ++r)
      if (r != rank) {
        if (ctx->peers.x[r]) cudaIpcCloseMemHandle(ctx->peers.x[r]);
        if (ctx->peers.flags[r]) cudaIpcCloseMemHandle(ctx->peers.flags[r]);
        ctx->peers.x[r] = nullptr; ctx->peers.flags[r] = nullptr;
      }
    (void)cudaGetLastError();
    return LQRO_OK;
  }
  ctx->p2p = true;
  ctx->p2p_epoch = ctx->p2p_done_epoch = 0;
  // blocks of the push kernel: 16 B per thread and peer, a few stores in flight per thread
  const size_t cnt2 = 8 * (size_t)ctx->prm.n_local;
  int blocks = (int)((cnt2 + 256 * 4 - 1) / (256 * 4));
  ctx->p2p_blocks = blocks < 1 ? 1 : (blocks > 2 * ctx->sm_count ? 2 * ctx->sm_count : blocks);
  return LQRO_OK;
}

extern "C" int lqro_comm_init(lqro_ctx* ctx, int rank, int world, const void* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!load_nccl()) return fail(ctx, LQRO_E_NCCL, "libnccl not found");
  if ((size_t)ctx->prm.n_local * world != (size_t)ctx->prm.N || ctx->prm.local_begin != rank * ctx->prm.n_local)
    return fail(ctx, LQRO_E_ARG, "agents must be split in equal contiguous blocks, block r owned by rank r");
  CU(cudaSetDevice(ctx->cfg.device));
  lqro_ncclUniqueId id;
  memcpy(&id, id128, 128);
  int rc = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
  if (rc != 0) return fail(ctx, LQRO_E_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclCommInitRank failed");
  ctx->rank = rank; ctx->world = world;
  return p2p_setup(ctx);
}

extern "C" int lqro_comm_p2p_active(lqro_ctx* ctx) { return ctx && ctx->p2p ? 1 : 0; }

extern "C" int lqro_device_state(lqro_ctx* ctx, double** d_x) {
  if (!ctx || !d_x) return LQRO_E_ARG;
  *d_x = ctx->d_x;
  return LQRO_OK;
}

extern "C" int lqro_upload_local_state(lqro_ctx* ctx, const double* x_local) {
  if (!ctx || !x_local) return fail(ctx, LQRO_E_ARG, "bad argument");
  CU(cudaSetDevice(ctx->cfg.device));
  CU(cudaMemcpyAsync(ctx->d_x + 16 * (size_t)ctx->prm.local_begin, x_local, sizeof(double) * 16 * (size_t)ctx->prm.n_local,
                     cudaMemcpyHostToDevice, ctx->stream));
  return LQRO_OK;
}

extern "C" int lqro_allgather_state(lqro_ctx* ctx);
extern "C" int lqro_step_sharded(lqro_ctx* ctx, const double* x_local, const double* vpref, const int* nbr, double* newV) {
  if (!ctx || !x_local || !vpref || !newV) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_dynamics) return fail(ctx, LQRO_E_STATE, "set_dynamics first");
  if (!nbr && !ctx->device_knn && !ctx->have_nbr) return fail(ctx, LQRO_E_STATE, "no neighbour table: the first step needs nbr (or LQRO_FLAG_DEVICE_KNN)");
  if (nbr) ctx->have_nbr = true;
  CU(cudaSetDevice(ctx->cfg.device));
  const LqroParams& p = ctx->prm;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(ctx->d_x + 16 * (size_t)p.local_begin, x_local, sizeof(double) * 16 * (size_t)p.n_local, cudaMemcpyHostToDevice, st));
  int rc = upload_step_inputs(ctx, vpref, nbr, st);
  if (rc) return rc;
  if ((rc = lqro_allgather_state(ctx))) return rc;
  if (!nbr && ctx->device_knn && (rc = launch_knn(ctx, ctx->d_x, ctx->d_nbr, st))) return rc;
  rc = launch_step(ctx, ctx->d_x, ctx->d_vpref, ctx->d_nbr, ctx->d_newV, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(newV, ctx->d_newV, sizeof(double) * 3 * (size_t)p.n_local, cudaMemcpyDeviceToHost, st));
  int ctr9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, peer_err = 0;  // d_ctr[1..9], as in lqro_step
  CU(cudaMemcpyAsync(ctr9, ctx->d_ctr + 1, sizeof(ctr9), cudaMemcpyDeviceToHost, st));
  if (ctx->p2p) CU(cudaMemcpyAsync(&peer_err, ctx->d_p2p_flags + LQRO_P2P_FLAG_ERR, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (peer_err) return fail(ctx, LQRO_E_NCCL, "peer-memory all-gather timed out waiting for another rank");
  const int queued = ctr9[0];
  if (ctr9[8]) return fail(ctx, LQRO_E_CUDA, "the neighbour table did not arrive on the device while the pair kernel waited for it");
  if (queued > ctx->hull_cap) return fail(ctx, LQRO_E_LIMIT, "hull queue overflow: inside pairs beyond max_hull_pairs got no half-plane (status 0x200)");
  return LQRO_OK;
}

extern "C" int lqro_allgather_state(lqro_ctx* ctx) {
  if (!ctx) return LQRO_E_ARG;
  if (!ctx->comm) return LQRO_OK;  // single GPU
  CU(cudaSetDevice(ctx->cfg.device));
  if (ctx->p2p) {
    // peer-memory push (lqro_p2p.cuh); a step that did not report "done reading" yet (lqro_allgather_state called on its own) does so now
    if (ctx->p2p_done_epoch < ctx->p2p_epoch) {
      k_p2p_done<<<1, 32, 0, ctx->stream>>>(ctx->peers, ctx->rank, ctx->world, ctx->p2p_epoch);
      ctx->p2p_done_epoch = ctx->p2p_epoch;
    }
    ++ctx->p2p_epoch;
    k_p2p_gather<<<ctx->p2p_blocks, 256, 0, ctx->stream>>>(ctx->peers, ctx->rank, ctx->world, 8 * (size_t)ctx->prm.n_local, ctx->p2p_epoch);
    CU(cudaGetLastError());
    return LQRO_OK;
  }
  const size_t cnt = 16 * (size_t)ctx->prm.n_local;
  const int ncclFloat64 = 8;
  int rc = g_nccl.AllGather(ctx->d_x + cnt * ctx->rank, ctx->d_x, cnt, ncclFloat64, ctx->comm, ctx->stream);
  if (rc != 0) return fail(ctx, LQRO_E_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclAllGather failed");
  return LQRO_OK;
}

extern "C" int lqro_sim_init(lqro_ctx* ctx, int n, const double* x, const double* Rot, const double* P, const double* pGoal,
                             const double* normals16, double* xTrue, double* RotTrue, double* vGoal) {
  if (!ctx || n < 0 || !x || !Rot || !P || !pGoal || !normals16 || !xTrue || !RotTrue || !vGoal) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_sim_model) return fail(ctx, LQRO_E_STATE, "lqro_sim_set_model first");
  if (n == 0) return LQRO_OK;
  CU(cudaSetDevice(ctx->cfg.device));
  const size_t N = (size_t)n, per = 16 + 9 + 256 + 3 + 16 + 16 + 9 + 3;
  int rc = ensure_scratch(ctx, sizeof(double) * per * N);
  if (rc) return rc;
  double *dx = (double*)ctx->d_scratch, *dR = dx + 16 * N, *dP = dR + 9 * N, *dpg = dP + 256 * N, *dn = dpg + 3 * N, *dxT = dn + 16 * N,
         *dRT = dxT + 16 * N, *dvG = dRT + 9 * N;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(dx, x, 8 * 16 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dR, Rot, 8 * 9 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dP, P, 8 * 256 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dpg, pGoal, 8 * 3 * N, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dn, normals16, 8 * 16 * N, cudaMemcpyHostToDevice, st));
  const int wpb = 4;
  int blocks = (n + wpb - 1) / wpb;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  k_sim_init<<<blocks, wpb * 32, wpb * sizeof(LqroSimWsB), st>>>(n, ctx->d_sim_model, dx, dR, dP, dpg, dn, dxT, dRT, dvG);
  CU(cudaGetLastError());
  ctx->sim_launches++;
  CU(cudaMemcpyAsync(xTrue, dxT, 8 * 16 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(RotTrue, dRT, 8 * 9 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(vGoal, dvG, 8 * 3 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LQRO_OK;
}

extern "C" int lqro_linearize_discretize(lqro_ctx* ctx, int n_batch, const double* x, const double* Rot, const double* u, double* A, double* B,
                                         double* c) {
  if (!ctx || n_batch < 0 || !x || !Rot || !u || !A || !B || !c) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_sim_model) return fail(ctx, LQRO_E_STATE, "lqro_sim_set_model first");
  if (n_batch == 0) return LQRO_OK;
  CU(cudaSetDevice(ctx->cfg.device));
  const size_t nb = (size_t)n_batch, per = 16 + 9 + 4 + 256 + 64 + 16;
  int rc = ensure_scratch(ctx, sizeof(double) * per * nb);
  if (rc) return rc;
  double *dx = (double*)ctx->d_scratch, *dR = dx + 16 * nb, *du = dR + 9 * nb, *dA = du + 4 * nb, *dB = dA + 256 * nb, *dc = dB + 64 * nb;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(dx, x, 8 * 16 * nb, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dR, Rot, 8 * 9 * nb, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(du, u, 8 * 4 * nb, cudaMemcpyHostToDevice, st));
  const int wa = SIM_FLOW_WARPS;
  const size_t smem = (size_t)wa * sizeof(LqroSimWsA);
  CU(cudaFuncSetAttribute(k_sim_linearize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = (n_batch + wa - 1) / wa;
  if (blocks > ctx->sm_count * SIM_FLOW_CTAS) blocks = ctx->sm_count * SIM_FLOW_CTAS;
  k_sim_linearize<<<blocks, wa * 32, smem, st>>>(n_batch, ctx->d_sim_model, dx, dR, du, dA, dB, dc, ctx->sim_m_diag);
  CU(cudaGetLastError());
  ctx->sim_launches++;
  CU(cudaMemcpyAsync(A, dA, 8 * 256 * nb, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(B, dB, 8 * 64 * nb, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c, dc, 8 * 16 * nb, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return LQRO_OK;
}

// One whole time step of the reference's loop (MAIN:1373-1417) with everything resident: [all-gather of the estimates] ->
// [neighbour table] -> avoidance step -> simulator-form update of the owned agents.  Launches only; no host synchronisation.
extern "C" int lqro_timestep_device(lqro_ctx* ctx, double* d_Rot, double* d_P, double* d_xTrue, double* d_RotTrue, const double* d_pGoal,
                                    const double* d_normals, double* d_vgoal, void* stream) {
  if (!ctx || !d_Rot || !d_P || !d_xTrue || !d_RotTrue || !d_pGoal || !d_normals || !d_vgoal) return fail(ctx, LQRO_E_ARG, "bad argument");
  if (!ctx->have_dynamics) return fail(ctx, LQRO_E_STATE, "set_dynamics first");
  if (!ctx->have_sim_model) return fail(ctx, LQRO_E_STATE, "lqro_sim_set_model first");
  CU(cudaSetDevice(ctx->cfg.device));
  const LqroParams& p = ctx->prm;
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  int rc = LQRO_OK;
  if (ctx->comm) {
    if (st != ctx->stream) return fail(ctx, LQRO_E_ARG, "with a communicator the time step runs on the context's stream (lqro_set_stream)");
    if ((rc = lqro_allgather_state(ctx))) return rc;
  }
  ctx->h_nbr_chunked = nullptr; ctx->h_nbr_streamed = nullptr; ctx->h_vpref_streamed = nullptr;
  if (!ctx->device_knn && !ctx->have_nbr) return fail(ctx, LQRO_E_STATE, "no neighbour table: step once with a host table, lqro_knn_device, or LQRO_FLAG_DEVICE_KNN");
  if (ctx->device_knn && (rc = launch_knn(ctx, ctx->d_x, ctx->d_nbr, st))) return rc;
  if ((rc = launch_step(ctx, ctx->d_x, d_vgoal, ctx->d_nbr, ctx->d_newV, st))) return rc;
  return launch_sim_update(ctx, p.n_local, ctx->d_x + 16 * (size_t)p.local_begin, d_Rot, d_P, d_xTrue, d_RotTrue, ctx->d_newV, d_pGoal, d_normals,
                           d_vgoal, nullptr, st);
}

// ---- measurement helpers ------------------------------------------------------------------------------
template <typename T, typename Kern>
static int measure_peak(lqro_ctx* ctx, Kern kern, double* tflops) {
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  int rc = ensure_scratch(ctx, sizeof(T) * (size_t)blocks * threads);
  if (rc) return rc;
  T* d = (T*)ctx->d_scratch;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    kern<<<blocks, threads, 0, ctx->stream>>>(d, iters);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaEventSynchronize(ctx->ev[1]));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    if (rep > 0 && ms < best) best = ms;
  }
  const double flop = 2.0 * 8.0 * iters * (double)blocks * threads;
  *tflops = flop / (best * 1e-3) / 1e12;
  return LQRO_OK;
}
extern "C" int lqro_measure_fp64_peak(lqro_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  return measure_peak<double>(ctx, k_peak_fp64, tflops);
}
extern "C" int lqro_measure_fp32_peak(lqro_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return LQRO_E_ARG;
  CU(cudaSetDevice(ctx->cfg.device));
  return measure_peak<float>(ctx, k_peak_fp32, tflops);
}
