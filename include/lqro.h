/* include/lqro.h -- C ABI of the B200-native LQR-Obstacles hot path (liblqro.so).
 *
 * The reference (hihixuyang/LQR-Obstacles) has no plugin/FFI layer: its boundary is the set of free
 * functions its main loop calls every time step.  Each entry point below names the reference
 * function(s) it replaces (aliases as in SURVEY.md: MAIN = QuadrotorHoverController/
 * LQRObstaclesPositionControl(Working).cpp, AGENT = QuadrotorHoverController/Agent.cpp).
 * The C++ shims carrying the reference's own signatures are in include/lqro_dropin.hpp and call
 * only what is declared here.
 *
 * Conventions: every function returns 0 on success and a negative LQRO_E_* code on failure; nothing
 * throws or aborts.  The caller owns every host buffer; the context owns all device memory, streams,
 * events and the NCCL communicator.  One context per GPU per process; calls on one context must be
 * serialised by the caller (the reference is single-threaded), different contexts are independent.
 * There is NO CPU fallback: without a CUDA device lqro_create fails with LQRO_E_CUDA.
 *
 * Matrices are row-major FP64.  An agent's state is the reference's 16-vector
 * [p(3) v(3) r(3) w(3) rotor forces(4)] (MAIN:341-345, simulator2.h:4).
 */
#ifndef LQRO_H
#define LQRO_H

#ifdef __cplusplus
extern "C" {
#endif

#define LQRO_X_DIM 16
#define LQRO_U_DIM 4
#define LQRO_V_DIM 3

#define LQRO_OK            0
#define LQRO_E_ARG        -1   /* bad argument */
#define LQRO_E_CUDA       -2   /* CUDA runtime error / no device (see lqro_last_error) */
#define LQRO_E_STATE      -3   /* call order (e.g. step before set_dynamics) */
#define LQRO_E_NCCL       -4   /* NCCL not available / communicator error */
#define LQRO_E_LIMIT      -5   /* a configured capacity was exceeded */

/* hull_input: what the convex hull is computed on (MAIN:765-767 prints the cloud with 6
 * significant digits before qconvex reads it). */
#define LQRO_HULL_TEXT6    0   /* reference behaviour: strtod(printf("%g")) of every coordinate */
#define LQRO_HULL_FP64     1   /* un-rounded FP64 cloud */

/* per-pair status written by a step */
#define LQRO_PAIR_NONE     0   /* <= 4 reachable points or empty slot: no plane (MAIN:1393) */
#define LQRO_PAIR_OUTSIDE  1   /* relative velocity outside the obstacle: GJK plane (MAIN:1399) */
#define LQRO_PAIR_INSIDE   2   /* inside: nearest-hull-facet plane (MAIN:1396) */
#define LQRO_PAIR_DEGENERATE 0x100 /* OR-ed in: the hull met a round-off-band decision / internal limit */
#define LQRO_PAIR_DROPPED  0x200   /* inside pair that did not fit the hull work queue (max_hull_pairs): carries NO plane
                                      (low byte 0), counted in lqro_get_counters()[4]; enlarge max_hull_pairs */

typedef struct lqro_ctx lqro_ctx;

/* One POD carries what the reference keeps in #defines and globals (MAIN:30-36, 1102, 1367). */
typedef struct lqro_config {
  int device;            /* CUDA device ordinal */
  int n_agents;          /* N: agents whose state is visible to this context (global count) */
  int local_begin;       /* first agent owned by this context (0 on one GPU) */
  int n_local;           /* agents owned (0 => n_agents - local_begin) */
  int max_neighbors;     /* K: neighbour slots per agent (reference: N-1, every other agent) */
  int horizon;           /* H = OBSTACLE_STEPS (30) */
  int n_samples;         /* P = NUM_POINTS (50) */
  double xy_radius;      /* XYRADIUS 0.26 */
  double z_radius;       /* ZRADIUS 0.75 */
  double max_speed;      /* reachable-velocity sphere radius, MAIN:1367 (30) */
  double lp_radius;      /* maxSpeed_ of calculateNewV, MAIN:1102 (100) */
  int hull_input;        /* LQRO_HULL_TEXT6 | LQRO_HULL_FP64 */
  int fix_facet0_normal; /* 1: nearest facet 0 yields facet 0's normal; 0 is not offered on the device
                            (the reference then keeps the previous pair's normal, a sequential
                            cross-pair dependency) */
  int max_hull_pairs;    /* capacity of the inside-case work queue per step (0 => n_local*K) */
  int hull_slots;        /* concurrent hull workspaces (0 => default) */
  int flags;             /* LQRO_FLAG_* (0 = defaults) */
} lqro_config;

#define LQRO_FLAG_NO_PRUNE 1   /* evaluate every horizon step of every pair (ablation / parity of the exact pruning) */
#define LQRO_FLAG_DEVICE_KNN 4 /* lqro_step / lqro_step_sharded called with nbr == NULL rebuild the neighbour table on the
                                  device from the current positions every step (lqro_knn_device) instead of reusing the
                                  table uploaded last */
#define LQRO_FLAG_FUSED_HULL 2 /* one hull warp per pair-kernel CTA consumes the inside-case queue concurrently
                                  (experimental: measured slower than the stand-alone hull kernel) */

#define LQRO_FLAG_NO_OVERLAP 8 /* steps of large swarms (>= 2^17 pairs) run a polling hull launch on a few SMs of its own next to the
                                  pair kernel, as a concurrent consumer of the inside-case queue; this flag (or LQRO_HULL_OVERLAP_SMS=0
                                  in the environment) keeps the pair kernel on every SM followed by the hulls.  Host-buffer steps
                                  then upload the neighbour table in blocks overlapped with per-block pair launches. */

/* Fills the reference's compile-time constants (MAIN:30-36, 1102, 1367). */
void lqro_default_config(lqro_config* cfg);

int lqro_create(const lqro_config* cfg, lqro_ctx** out);
int lqro_destroy(lqro_ctx* ctx);
const char* lqro_last_error(const lqro_ctx* ctx);  /* ctx may be NULL: error of the last failed lqro_create */
const char* lqro_version(void);

/* ---- (1) dynamics / gains ------------------------------------------------------------------- */

/* Replaces the Riccati part of controlMatrices (MAIN:512-543): for each of n_batch problems the
 * velocity-LQR gains L (4x16), E (4x3) and the position-LQR gains Lh (3x16), Eh (3x3) of the
 * linearised model A (16x16), B (16x4) with weights Qv = qv*I3, Qp = qp*I3, R = r*I4, Qx = 0, c = 0,
 * by `iters` fixed-point iterations on the device (warp per problem).  Host pointers. */
int lqro_solve_gains(lqro_ctx* ctx, int n_batch, const double* A, const double* B, double qv, double qp, double r,
                     int iters, double* L, double* E, double* Lh, double* Eh);
/* Device time (CUDA events around the launch) of the Riccati kernel of the last lqro_solve_gains, in milliseconds. */
int lqro_riccati_last_ms(lqro_ctx* ctx, float* ms);

/* Replaces findFG (MAIN:653-662) x H and the pair-independent half of createObstacle (MAIN:684-685):
 * builds on the device, in the reference's operation order, NF_k = -(C F_k) and T_k = (C G_k)^-1 for
 * k = 1..H with F_k = (A+BL)^k, G_k = sum_{m<k} (A+BL)^m B E.  Must precede any step. Host pointers. */
int lqro_set_dynamics(lqro_ctx* ctx, const double* A, const double* B, const double* L, const double* E);

/* Heterogeneous dynamics (dynamics classes, e.g. mass / inertia / latency variants): the same for class `cls` in
 * [0, 16); class 0 is what lqro_set_dynamics sets.  The obstacle of a pair (i, j) is built from the matrices of agent
 * i (findFG(A, B, qlist[i]->L, qlist[i]->E, ...), MAIN:1384), i.e. from the tables of i's class. */
int lqro_set_dynamics_class(lqro_ctx* ctx, int cls, const double* A, const double* B, const double* L, const double* E);
/* Class of every agent: n_agents ints (host), every class named must have its dynamics set.  NULL: all class 0.
 * The step then runs the pair kernel once per class over the owned agents of that class. */
int lqro_set_agent_classes(lqro_ctx* ctx, const int* cls);

/* createSpheres (MAIN:665-680): the P samples of the doubled ellipsoid (P*3 doubles), available right after lqro_create. */
int lqro_get_spheres(lqro_ctx* ctx, double* sph);
/* Copies the tables back (parity checks): NF H*3*16, T H*3*3, sphere samples P*3 (createSpheres,
 * MAIN:665-680).  Any pointer may be NULL. */
int lqro_get_tables(lqro_ctx* ctx, double* NF, double* T, double* spheres);

/* ---- (2)-(4) the per-timestep path ------------------------------------------------------------ */

/* One avoidance step for the agents this context owns: for every owned agent i and neighbour slot
 * kk (j = nbr[i_local*K + kk], j < 0 = empty) createObstacle x H, findReachableObstacle, pointInHull /
 * run_gjk / convexHull, createHalfPlanes (MAIN:1381-1403), then calculateNewV (MAIN:1406).
 * HOST buffers; the copies to and from the device are part of the call.
 *   x      n_agents*16  state ESTIMATES of all agents (Quadrotor::x, MAIN:1385)
 *   vpref  n_local*3    preferred velocities (Quadrotor::vGoal)
 *   nbr    n_local*K    neighbour agent ids (global), planes are generated in slot order
 *   newV   n_local*3    out: Quadrotor::newV */
int lqro_step(lqro_ctx* ctx, const double* x, const double* vpref, const int* nbr, double* newV);

/* Same with DEVICE pointers, asynchronous on the context's stream (or `stream` if non-NULL, a
 * cudaStream_t).  d_x may be NULL to use the context's own state buffer (see lqro_device_state). */
int lqro_step_device(lqro_ctx* ctx, const double* d_x, const double* d_vpref, const int* d_nbr, double* d_newV,
                     void* stream);
int lqro_synchronize(lqro_ctx* ctx);
/* Make the context launch on a caller-owned cudaStream_t (e.g. the harness's stream, so its events see
 * the kernels).  NULL restores the context's own stream. */
int lqro_set_stream(lqro_ctx* ctx, void* stream);

/* Per-pair results of the last step (host pointers, any may be NULL):
 *   planes n_local*K*6 float (point xyz, normal xyz; createHalfPlanes, MAIN:1085-1098)
 *   status n_local*K   int   (LQRO_PAIR_*)
 *   nreach n_local*K   int   (points surviving findReachableObstacle)
 *   dist   n_local*K   double (distance before the 0.5 of MAIN:1401)
 *   nfacets n_local*K  int   (hull facets, inside pairs only) */
int lqro_get_pair_results(lqro_ctx* ctx, float* planes, int* status, int* nreach, double* dist, int* nfacets);

/* counters of the last step: [0] pairs evaluated, [1] pairs with a plane, [2] inside (hull) pairs,
 * [3] degenerate hulls, [4] hull queue overflow, [5] kernels launched by the step */
int lqro_get_counters(lqro_ctx* ctx, long long* out6);

/* ---- per-stage entry points (unit parity; host pointers) -------------------------------------- */

/* findFG (MAIN:653-662), one call with the reference's own meaning: F <- (A + B L) F, G <- (A + B L) G + B E, in place
 * (A 16x16, B 16x4, L 4x16, E 4x3, F 16x16, G 16x3, row major), every element in the reference's operation order: the
 * result has the reference's bits.  The step itself never calls this per pair: F_k, G_k are pair independent and tabulated
 * once by lqro_set_dynamics. */
int lqro_find_fg(lqro_ctx* ctx, const double* A, const double* B, const double* L, const double* E, double* F, double* G);
/* createObstacle (MAIN:683-688), one call for one horizon step with the caller's F, G and position map P (3x16):
 * Transform = !(P G) (3x3), Translate = -P F (x1 - x2) (3), q[s] = Transform (points[s] + Translate), s < n_points. */
int lqro_create_obstacle(lqro_ctx* ctx, int n_points, const double* points, const double* F, const double* G, const double* P,
                         const double* x1, const double* x2, double* Transform, double* Translate, double* q);
/* createObstacle x H for one pair (MAIN:683-688): q = H*P*3 candidate points, k-major. */
int lqro_candidates(lqro_ctx* ctx, const double* xi, const double* xj, double* q);
/* findReachableObstacle (MAIN:691-706): ascending indices of reachable points; returns count in *n_out. */
int lqro_reachable(lqro_ctx* ctx, const double* q, int n, const double* xi, const double* xj, int* idx, int* n_out);
/* run_gjk (MAIN:712-746): squared distance, the two witness points (on {v_rel} -- a rounded convex combination
 * of v_rel with itself, GJK.c:851-862 -- and on the cloud), G-test count.  Optional outputs may be NULL. */
int lqro_gjk(lqro_ctx* ctx, const double* vrel, const double* pts, int n, double* sqdist, double* wpt_vrel, double* wpt_hull,
             int* iterations);
/* qconvex n + Fv (MAIN:762-806) for one cloud given as-is (no rounding applied here): facets in qhull
 * order, planes nf*4, verts nf*3 (first vertex = apex).  *nf_io: in capacity, out count. */
int lqro_convex_hull(lqro_ctx* ctx, const double* pts, int n, double* planes, int* verts, int* nf_io, int* degenerate);
/* convexHull (MAIN:809-838): 6-digit rounding per hull_input, hull, nearest facet. */
int lqro_hull_escape(lqro_ctx* ctx, const double* pts, int n, const double* vrel, double* distance, double* normal,
                     int* best_facet, int* n_facets);
/* calculateNewV (MAIN:1101-1112) for `batch` independent agents: planes batch*max_planes*6 floats,
 * counts batch, goal batch*3 FP64, newV batch*3 FP64. */
int lqro_new_velocity(lqro_ctx* ctx, int batch, int max_planes, const float* planes, const int* counts,
                      const double* goal, double* newV);
/* RVO-form Agent::update (AGENT:176-180), FP32: vel = newVel; pos += vel*dt.  Host pointers, n*3 each. */
int lqro_agent_update(lqro_ctx* ctx, int n, float* pos, float* vel, const float* newVel, float dt);
/* Neighbour search (the list RVO's Agent::insertAgentNeighbor builds, AGENT:153-174; the KdTree that feeds it is
 * not in the reference): for every owned agent the max_neighbors other agents with the smallest squared
 * distance (dx*dx + dy*dy) + dz*dz, ascending, equal distances to the lower index, -1 = none.  Uniform grid on
 * the device.  x n_agents*16 (positions = entries 0..2; NULL = the state already resident), nbr n_local*K out.
 * The table also stays resident as the context's neighbour table. */
int lqro_knn(lqro_ctx* ctx, const double* x, int* nbr);
/* Same with device pointers, asynchronous (d_x NULL = resident state, d_nbr NULL = the resident table). */
int lqro_knn_device(lqro_ctx* ctx, const double* d_x, int* d_nbr, void* stream);
/* Copies the resident neighbour table (n_local*K) to the host. */
int lqro_get_neighbors(lqro_ctx* ctx, int* nbr);
/* printf("%g")/strtod round trip of MAIN:765-767 applied to n doubles on the device. */
int lqro_round6(lqro_ctx* ctx, const double* in, int n, double* out);

/* ---- simulator-form agent update: the second loop of the reference's time step ------------ */

/* What the reference keeps in globals (MAIN:40-76; values set in setup() MAIN:176-198 and _tmain MAIN:1126-1140) and in
 * every Quadrotor (gains, MAIN:95-99).  Row-major FP64. */
typedef struct lqro_sim_model {
  double dt, gravity, mass, moment_const, thrust_latency, length, jstep;
  double inertia[9], inv_inertia[9];
  double u_goal[4];                              /* uGoal = nominalInput x 4 (MAIN:1126-1127) */
  double L[64], E[12], l[4], Lh[48], Eh[9];      /* controlMatrices outputs (MAIN:490-544) */
  double M[256], N[36];                          /* motion / observation noise covariances (MAIN:1139-1140) */
} lqro_sim_model;
/* Fills the reference's constants (setup(), MAIN:176-198; M = N = 1e-9 I); gains are zeroed: copy L, E, l, Lh, Eh in
 * (from lqro_solve_gains or the caller's controlMatrices). */
void lqro_sim_default_model(lqro_sim_model* m);
int lqro_sim_set_model(lqro_ctx* ctx, const lqro_sim_model* m);

/* Replaces, for n agents at once, the loop body MAIN:1409-1416:
 *     vGoal = newV; u = findU(); propagateU(u); kalmanFilter1(x, Rot, u, M, P);
 *     z = sampleGaussian(h(xTrue, RotTrue), N); kalmanFilter2(x, Rot, z, N, P); vGoal = findVGoal();
 * (riccatiControllerSteady MAIN:547-576, propagate MAIN:442-455, kalmanFilter1/2 MAIN:457-487,
 * riccatiControllerSteadyPosition MAIN:579-608), one warp per agent (per flow) in each of the update's kernels.  HOST pointers, updated in place:
 *   x n*16 estimate, Rot n*9, P n*256 covariance, xTrue n*16, RotTrue n*9 (Quadrotor::x, Rot, P, xTrue, RotTrue)
 *   newV n*3 (calculateNewV's output), pGoal n*3 (Quadrotor::pGoal)
 *   normals n*22: the standard normals the reference's sampleGaussian calls draw for agent i, in its order -- 16 for
 *           propagate's process noise (MAIN:450), then 6 for the observation (MAIN:1413); see lqro_msvc_normals
 *   vGoal n*3 out (next preferred velocity), u n*4 out (control input; may be NULL) */
int lqro_sim_update(lqro_ctx* ctx, int n, double* x, double* Rot, double* P, double* xTrue, double* RotTrue,
                    const double* newV, const double* pGoal, const double* normals, double* vGoal, double* u);
/* Same with DEVICE pointers, asynchronous on the context's stream (or `stream`).  d_x may be the context's own state
 * buffer (lqro_device_state), d_newV the d_newV of lqro_step_device and d_vGoal its d_vpref for the next step, so that a
 * whole time step stays on the device. */
int lqro_sim_update_device(lqro_ctx* ctx, int n, double* d_x, double* d_Rot, double* d_P, double* d_xTrue, double* d_RotTrue,
                           const double* d_newV, const double* d_pGoal, const double* d_normals, double* d_vGoal, double* d_u,
                           void* stream);
/* Replaces the noise part of Quadrotor::setupQuadrotors (MAIN:126-129) and the first findVGoal (MAIN:1355) for n agents:
 * xTrue = sampleGaussian(x, P); RotTrue = Rot*exp([xTrue_r]); xTrue_r = 0; vGoal = findVGoal().  Host pointers:
 * x n*16, Rot n*9, P n*256, pGoal n*3, normals16 n*16 (the 16 draws of normal() per agent, in agent order) in;
 * xTrue n*16, RotTrue n*9, vGoal n*3 out. */
int lqro_sim_init(lqro_ctx* ctx, int n, const double* x, const double* Rot, const double* P, const double* pGoal,
                  const double* normals16, double* xTrue, double* RotTrue, double* vGoal);
/* Replaces linearizeDiscretize (MAIN:425-440; Jacobian_fx / Jacobian_fu MAIN:392-412, exp MAT:773-790), batched over
 * n_batch operating points of the model set by lqro_sim_set_model (its gains are not used): A = exp(dt F) 16x16,
 * B = Int*G 16x4, c = Int*f 16, Int = (dt/6)(I + 4 exp(dt/2 F) + A).  Followed by lqro_solve_gains this is the whole of
 * controlMatrices (MAIN:490-544) on the device: x = [0 .. 0, uGoal], Rot = I, u = uGoal there.  Host pointers:
 * x n*16, Rot n*9, u n*4 in; A n*256, B n*64, c n*16 out. */
int lqro_linearize_discretize(lqro_ctx* ctx, int n_batch, const double* x, const double* Rot, const double* u, double* A, double* B,
                              double* c);
/* One WHOLE time step of the reference's loop (MAIN:1373-1417) for the owned agents with everything resident on the device,
 * launches only: [ncclAllGather of the estimates when a communicator exists] -> [neighbour table rebuilt from the current
 * positions with LQRO_FLAG_DEVICE_KNN, else the resident table] -> the avoidance step (as lqro_step_device on the context's
 * state buffer, lqro_device_state) -> the simulator-form update of the owned agents' block of that buffer.
 *   d_Rot, d_P, d_xTrue, d_RotTrue, d_pGoal, d_normals: as lqro_sim_update_device, n_local agents
 *   d_vgoal n_local*3  in: Quadrotor::vGoal (the preferred velocities of this step), out: the next step's
 * Per step only the 22*n_local standard normals have to come from the host. */
int lqro_timestep_device(lqro_ctx* ctx, double* d_Rot, double* d_P, double* d_xTrue, double* d_RotTrue, const double* d_pGoal,
                         const double* d_normals, double* d_vgoal, void* stream);
/* Host helper, no device work: `count` draws of the reference's normal() (MAIN:325-336) from the MSVC CRT rand() stream
 * it was written against (RAND_MAX 32767, seed*214013+2531011; random() MAIN:320-322).  *seed_io: generator state in,
 * state out (srand(1) at program start => 1).  The stream is sequential across agents, which is why the noise is an
 * input of lqro_sim_update rather than generated on the device. */
int lqro_msvc_normals(unsigned int* seed_io, int count, double* out);
/* Device time (ms) of the last lqro_sim_update / lqro_sim_update_device (CUDA events on the launching stream;
 * synchronises): [0] all kernels together, [1] k_sim_control + k_sim_flow (control input, the two linearised flows of every
 * agent), [2] k_sim_noise (process-noise sample), [3] k_sim_filter (state / covariance updates of kalmanFilter1 and propagate,
 * observation, kalmanFilter2, findVGoal). */
int lqro_sim_last_ms(lqro_ctx* ctx, float* out4);

/* ---- multi-GPU: agents partitioned, one all-gather of the state per step --------------------- */

/* 128-byte NCCL unique id (rank 0 creates it, the caller ships it to the other ranks). */
int lqro_comm_unique_id(void* id128);
int lqro_comm_init(lqro_ctx* ctx, int rank, int world, const void* id128);
/* 1 when the all-gather of the sharded step runs over NVLink peer memory (every rank pushes its block of the state straight into
 * every peer's state buffer, csrc/lqro_p2p.cuh; set up inside lqro_comm_init through CUDA IPC handles exchanged over the new
 * communicator, used only if EVERY rank could map every peer), 0 when it is ncclAllGather (fallback; LQRO_NO_P2P_GATHER=1 forces it). */
int lqro_comm_p2p_active(lqro_ctx* ctx);
/* Device pointer to the context's n_agents*16 state buffer; each rank writes its own block
 * [local_begin, local_begin+n_local) there (lqro_upload_local_state or its own kernels). */
int lqro_device_state(lqro_ctx* ctx, double** d_x);
int lqro_upload_local_state(lqro_ctx* ctx, const double* x_local);
/* The sharded step with HOST buffers (multi-GPU form of lqro_step): uploads this rank's block of the
 * state (x_local n_local*16), all-gathers it, runs the step for the owned agents, downloads newV.
 * With no communicator it equals lqro_step (n_local == n_agents). */
int lqro_step_sharded(lqro_ctx* ctx, const double* x_local, const double* vpref, const int* nbr, double* newV);
/* ncclAllGather of the per-rank blocks into every rank's state buffer (in place), on the context's stream.
 * With no communicator (single GPU) it is a no-op. Blocks must be equal-sized (n_local identical). */
int lqro_allgather_state(lqro_ctx* ctx);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* Peak-rate micro-benchmarks used for the roofline denominators that MEASURED_PEAKS.json lacks:
 * chains of independent DFMA / FFMA on every SM; returns TFLOP/s (2 flop per FMA). */
int lqro_measure_fp64_peak(lqro_ctx* ctx, double* tflops);
int lqro_measure_fp32_peak(lqro_ctx* ctx, double* tflops);
/* Device time (ms) of the kernels of the last step: [0] pair kernel, [1] hull kernel, [2] LP kernel,
 * [3] whole step (kernels only, measured with CUDA events on the launching stream). */
int lqro_last_step_ms(lqro_ctx* ctx, float* out4);
/* Steps of large swarms (polling hull launch next to the pair kernel): [0] set-up before the pair kernel, [1] the pair kernel (on
 * all but [3] SMs), [2] the polling hull launch (its own stream, started together with the pair kernel), [3] its SMs, [4] neighbour
 * slots per agent in the first chunk of the slot-major enumeration; other steps: [1] only ([3] = 0). */
int lqro_last_step_detail_ms(lqro_ctx* ctx, float* out5);
/* 1 if the host neighbour table (and the preferred velocities) of the last lqro_step / lqro_step_sharded went up NEXT TO the pair
 * kernel -- rows in two blocks on a copy stream, the kernel waiting for the rows of a batch -- instead of in front of it: steps
 * of >= 2^21 pairs with the polling hull launch (LQRO_NBR_STREAM_MIN_PAIRS=<pairs> at lqro_create moves the threshold, 0 = always
 * copy in front).  Results are
 * the same bits either way (the pairs are enumerated in another order, nothing else).  Replaces the reference's in-place
 * neighbour lists (Agent.cpp:60-82 fill them, MAIN:1379 walks them). */
int lqro_last_step_nbr_streamed(lqro_ctx* ctx);
/* SM-cycle breakdown of the hull kernel of the last step, summed over hulls (development aid):
 * [0] init+simplex [1] partitionall [2] furthest+horizon+cone [3] planes+match [4] list [5] walk [6] redo
 * [7] append [8] delete [9] - [10] insertions [11] points re-partitioned [12] max cycles of one hull
 * [13] max points [14] max facets */
int lqro_debug_hull_cycles(lqro_ctx* ctx, long long* out16);
/* Profiling builds (-DLQRO_HULL_PROF): the same phases for the LONGEST hull of the last step; [12] its points, [13] its facets. */
int lqro_debug_hull_longest(lqro_ctx* ctx, long long* out16);
/* SM-cycle breakdown of the simulator-form update kernel, summed over agents since the last reset; zeros unless the
 * library was built with -DLQRO_SIM_PROF (tools/simprof.py).  Phases: see csrc/lqro_sim.cuh. */
int lqro_debug_sim_cycles(lqro_ctx* ctx, long long* out16, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LQRO_H */
