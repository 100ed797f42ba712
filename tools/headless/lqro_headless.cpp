// tools/headless/lqro_headless.cpp -- the reference's simulator main loop (_tmain, MAIN:1114-1423) headless on Linux,
// written against the C ABI only (include/lqro.h): plain C++ host code, every computation on the GPU.
//
//   setup (MAIN:1126-1356)   constants (lqro_sim_default_model) -> controlMatrices = lqro_linearize_discretize at the hover
//                            point + lqro_solve_gains -> lqro_set_dynamics (findFG for the whole horizon) -> initial true
//                            states and goal velocities (lqro_sim_init)
//   time loop (MAIN:1371-1418)  lqro_step (obstacles, half-planes, new velocities) + lqro_sim_update (LQR feedback, propagate,
//                            Kalman filter, next goal velocity); the Gaussian noise is the reference's normal() over the MSVC
//                            rand() stream (lqro_msvc_normals), consumed in the reference's order
//   visualisation            none (Callisto): the trajectory goes to a binary dump instead
//
// usage: lqro_headless [--scenario hexagon6 | circle:N:RADIUS | file:PATH] [--steps T] [--seed S] [--horizon H]
//                      [--out traj.bin] [--device D] [--quiet]
//   scenario file: first line N, then N lines "sx sy sz gx gy gz" (start and goal positions, metres)
//   hexagon6 = the canonical scenario of the reference (MAIN:1163-1218); circle = N quads on a circle, antipodal goals
// trajectory dump (little endian): 8 bytes magic "LQROTRJ1", int32 N, int32 steps, int32 16, int32 0, then per step
//   N*16 doubles xTrue (true states) followed by N*16 doubles x (EKF estimates), state layout [p v r w f] (simulator2.h:4)
// stdout: one JSON line of run metrics.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lqro.h"

static void die(lqro_ctx* ctx, int rc, const char* what) {
  std::fprintf(stderr, "lqro_headless: %s failed (%d): %s\n", what, rc, lqro_last_error(ctx));
  std::exit(1);
}
#define CK(call) do { int rc__ = (call); if (rc__ != LQRO_OK) die(ctx, rc__, #call); } while (0)

int main(int argc, char** argv) {
  std::string scenario = "hexagon6", out;
  int steps = 210, horizon = 30, device = 0, quiet = 0;
  unsigned int seed = 1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&](const char* name) -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", name); std::exit(2); } return argv[++i]; };
    if (a == "--scenario") scenario = next("--scenario");
    else if (a == "--steps") steps = std::atoi(next("--steps"));
    else if (a == "--seed") seed = (unsigned int)std::strtoul(next("--seed"), nullptr, 10);
    else if (a == "--horizon") horizon = std::atoi(next("--horizon"));
    else if (a == "--out") out = next("--out");
    else if (a == "--device") device = std::atoi(next("--device"));
    else if (a == "--quiet") quiet = 1;
    else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  // ---- scenario: start and goal positions
  std::vector<double> start, goal;
  if (scenario == "hexagon6") {  // MAIN:1163-1218: hexagon of radius 6, goals at the antipodes
    const double d = std::sqrt(3.0);
    const double p[6][3] = {{6, 0, 0}, {3, 3 * d, 0}, {-3, 3 * d, 0}, {-6, 0, 0}, {-3, -3 * d, 0}, {3, -3 * d, 0}};
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 3; ++k) { start.push_back(p[i][k]); goal.push_back(p[(i + 3) % 6][k]); }
  } else if (scenario.rfind("circle:", 0) == 0) {
    int n = 0; double r = 0;
    if (std::sscanf(scenario.c_str(), "circle:%d:%lf", &n, &r) != 2 || n < 2) { std::fprintf(stderr, "circle:N:RADIUS\n"); return 2; }
    for (int i = 0; i < n; ++i) {
      const double a = 2.0 * M_PI * i / n;
      start.push_back(r * std::cos(a)); start.push_back(r * std::sin(a)); start.push_back(0.0);
      goal.push_back(-r * std::cos(a)); goal.push_back(-r * std::sin(a)); goal.push_back(0.0);
    }
  } else if (scenario.rfind("file:", 0) == 0) {
    FILE* f = std::fopen(scenario.c_str() + 5, "r");
    int n = 0;
    if (!f || std::fscanf(f, "%d", &n) != 1 || n < 2) { std::fprintf(stderr, "cannot read scenario %s\n", scenario.c_str() + 5); return 2; }
    start.resize(3 * (size_t)n); goal.resize(3 * (size_t)n);
    for (int i = 0; i < n; ++i)
      if (std::fscanf(f, "%lf %lf %lf %lf %lf %lf", &start[3 * i], &start[3 * i + 1], &start[3 * i + 2], &goal[3 * i], &goal[3 * i + 1], &goal[3 * i + 2]) != 6) {
        std::fprintf(stderr, "scenario line %d is short\n", i + 2); return 2;
      }
    std::fclose(f);
  } else { std::fprintf(stderr, "unknown scenario %s\n", scenario.c_str()); return 2; }
  const int n = (int)(start.size() / 3);
  const size_t N = (size_t)n;

  // ---- context: the reference's constants, every other quad as a neighbour (MAIN:1376-1380)
  lqro_config cfg;
  lqro_default_config(&cfg);
  cfg.device = device; cfg.n_agents = n; cfg.max_neighbors = n - 1; cfg.horizon = horizon;
  lqro_ctx* ctx = nullptr;
  CK(lqro_create(&cfg, &ctx));
  // ---- controlMatrices (MAIN:490-544) on the device
  lqro_sim_model model;
  lqro_sim_default_model(&model);
  CK(lqro_sim_set_model(ctx, &model));
  double xhat[16] = {0}, eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, A[256], B[64], c[16], L[64], E[12], Lh[48], Eh[9];
  for (int i = 0; i < 4; ++i) xhat[12 + i] = model.u_goal[i];
  CK(lqro_linearize_discretize(ctx, 1, xhat, eye, model.u_goal, A, B, c));
  CK(lqro_solve_gains(ctx, 1, A, B, 100.0, 0.1, 5.0, 300, L, E, Lh, Eh));  // Qv, Qp, R of MAIN:1130-1137
  std::memcpy(model.L, L, sizeof L); std::memcpy(model.E, E, sizeof E); std::memcpy(model.Lh, Lh, sizeof Lh); std::memcpy(model.Eh, Eh, sizeof Eh);
  CK(lqro_sim_set_model(ctx, &model));
  CK(lqro_set_dynamics(ctx, A, B, L, E));
  // ---- initial conditions (MAIN:1149-1161, setupQuadrotors MAIN:112-130)
  std::vector<double> x(16 * N, 0.0), xT(16 * N), R(9 * N), RT(9 * N), P(256 * N, 0.0), vGoal(3 * N), newV(3 * N), normals(22 * N);
  std::vector<int> nbr(N * (N - 1));
  for (size_t i = 0; i < N; ++i) {
    for (int k = 0; k < 3; ++k) x[16 * i + k] = start[3 * i + k];
    for (int k = 0; k < 4; ++k) x[16 * i + 12 + k] = model.u_goal[k];
    for (int k = 0; k < 9; ++k) R[9 * i + k] = eye[k];
    for (int k = 0; k < 16; ++k) P[256 * i + 17 * k] = 1.0 * 0.000000001;
    int s = 0;
    for (int j = 0; j < n; ++j) if (j != (int)i) nbr[i * (N - 1) + s++] = j;
  }
  CK(lqro_msvc_normals(&seed, 16 * n, normals.data()));
  CK(lqro_sim_init(ctx, n, x.data(), R.data(), P.data(), goal.data(), normals.data(), xT.data(), RT.data(), vGoal.data()));

  FILE* fo = nullptr;
  if (!out.empty()) {
    fo = std::fopen(out.c_str(), "wb");
    if (!fo) { std::fprintf(stderr, "cannot write %s\n", out.c_str()); return 2; }
    const int hdr[4] = {n, steps, 16, 0};
    std::fwrite("LQROTRJ1", 1, 8, fo);
    std::fwrite(hdr, sizeof(int), 4, fo);
  }
  // ---- time loop (MAIN:1371-1418)
  long long planes = 0, inside = 0;
  double min_sep = 1e300;
  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < steps; ++t) {
    CK(lqro_step(ctx, x.data(), vGoal.data(), nbr.data(), newV.data()));
    long long cnt[6];
    CK(lqro_get_counters(ctx, cnt));
    planes += cnt[1]; inside += cnt[2];
    CK(lqro_msvc_normals(&seed, 22 * n, normals.data()));
    CK(lqro_sim_update(ctx, n, x.data(), R.data(), P.data(), xT.data(), RT.data(), newV.data(), goal.data(), normals.data(), vGoal.data(), nullptr));
    if (fo) { std::fwrite(xT.data(), sizeof(double), 16 * N, fo); std::fwrite(x.data(), sizeof(double), 16 * N, fo); }
    for (size_t i = 0; i < N; ++i)
      for (size_t j = i + 1; j < N; ++j) {
        double d2 = 0;
        for (int k = 0; k < 3; ++k) { const double d = xT[16 * i + k] - xT[16 * j + k]; d2 += d * d; }
        if (d2 < min_sep) min_sep = d2;
      }
    if (!quiet && (t % 30 == 0 || t == steps - 1)) std::fprintf(stderr, "step %d\n", t);
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (fo) std::fclose(fo);
  double worst_goal = 0;
  for (size_t i = 0; i < N; ++i) {
    double d2 = 0;
    for (int k = 0; k < 3; ++k) { const double d = xT[16 * i + k] - goal[3 * i + k]; d2 += d * d; }
    if (d2 > worst_goal) worst_goal = d2;
  }
  std::printf("{\"scenario\": \"%s\", \"quads\": %d, \"steps\": %d, \"horizon\": %d, \"wall_s\": %.3f, \"ms_per_step\": %.3f, "
              "\"planes\": %lld, \"inside_pairs\": %lld, \"min_separation_m\": %.6f, \"max_distance_to_goal_m\": %.6f, \"rand_state\": %u}\n",
              scenario.c_str(), n, steps, horizon, wall, 1e3 * wall / (steps > 0 ? steps : 1), planes, inside, std::sqrt(min_sep),
              std::sqrt(worst_goal), seed);
  lqro_destroy(ctx);
  return 0;
}
