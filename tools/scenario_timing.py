#!/usr/bin/env python
"""BASELINE config 3: the 32- and 64-quadrotor reference scenarios (K = N-1, H = 30, P = 50) with a timing breakdown per
stage, device next to the reference's own CPU code (oracle/_ref; one core, as the reference is single-threaded).
The scenario is advanced by the reference simulator (LQR feedback, propagation, EKF) with the device step supplying
newV (bit-identical trajectory, tests/test_gpu_trajectory.py); every `stride`-th step is also timed on the CPU through
ref_step_agents.  Writes profiles/r2_scenarios_32_64.json and prints a markdown table."""
import ctypes, json, os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from _libs import Reference, RefSim
from lqr_obstacles_b200 import Lqro
from lqr_obstacles_b200.swarm import all_pairs_neighbors, circle_antipodal

ref = Reference()
ref.lib.ref_sim_set_fix_facet0(1)
out = []
for n, radius, steps, stride in ((32, 12.0, 60, 10), (64, 24.0, 60, 15)):
    S, G = circle_antipodal(n, radius)
    sim = RefSim(ref, S, G, seed=1)
    A, B, L, E, Lh, Eh = sim.dynamics()
    ctx = Lqro(n_agents=n, max_neighbors=n - 1, horizon=30)
    ctx.set_dynamics(A, B, L, E)
    nbr = all_pairs_neighbors(n)
    agents = np.arange(n, dtype=np.int32)
    g = {"pairs": [], "hull": [], "lp": [], "total": [], "wall": [], "inside": []}
    c = {"obstacle": [], "halfplane": [], "hull": [], "newv": [], "total": [], "hulls": []}
    for t in range(steps):
        x, xt, vg = sim.state()
        x = np.ascontiguousarray(x); vg = np.ascontiguousarray(vg)
        t0 = time.perf_counter()
        newV = ctx.step(x, vg, nbr)
        g["wall"].append((time.perf_counter() - t0) * 1e3)
        ms = ctx.last_step_ms()
        for k in ("pairs", "hull", "lp", "total"): g[k].append(ms[k])
        g["inside"].append(ctx.counters()["inside"])
        if t % stride == stride - 1:
            ref.lib.ref_stage_seconds(None, None, 1)
            t0 = time.perf_counter()
            ref.step_agents(A, B, L, E, x, vg, nbr, agents, 30, 30.0)
            c["total"].append((time.perf_counter() - t0) * 1e3)
            st = np.zeros(4); nh = ctypes.c_long()
            ref.lib.ref_stage_seconds(st.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nh), 1)
            c["obstacle"].append(st[0] * 1e3); c["halfplane"].append((st[1] - st[3]) * 1e3); c["hull"].append(st[3] * 1e3)
            c["newv"].append(st[2] * 1e3); c["hulls"].append(nh.value)
        sim.set_newv(newV); sim.update()
    sim.close(); ctx.close()
    m = lambda a: float(np.mean(a)) if len(a) else 0.0
    out.append({"quads": n, "pairs_per_step": n * (n - 1), "steps": steps, "H": 30, "P": 50,
                "device_ms": {"k_pairs": m(g["pairs"]), "k_hull": m(g["hull"]), "k_new_velocity": m(g["lp"]), "kernels": m(g["total"]),
                              "lqro_step_wall (host buffers)": m(g["wall"]), "inside_pairs_per_step": m(g["inside"]),
                              "k_hull_max": float(np.max(g["hull"]))},
                "reference_cpu_ms (1 core, in-process hull, no spawn)": {"obstacle_creation": m(c["obstacle"]), "half_plane_creation_without_hull": m(c["halfplane"]),
                                     "hull": m(c["hull"]), "new_velocity": m(c["newv"]), "step": m(c["total"]), "hulls_per_step": m(c["hulls"]),
                                     "steps_timed": len(c["total"])}})
json.dump(out, open(os.path.join(R, "profiles", "r2_scenarios_32_64.json"), "w"), indent=1)
print("| quads | pairs/step | inside pairs/step | k_pairs ms | k_hull ms | k_new_velocity ms | device step (kernels) ms | lqro_step wall ms | CPU obstacle ms | CPU half-plane ms | CPU hull ms | CPU new-V ms | CPU step ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for o in out:
    d, r = o["device_ms"], o["reference_cpu_ms (1 core, in-process hull, no spawn)"]
    print("| %d | %d | %.0f | %.3f | %.3f | %.3f | %.3f | %.3f | %.1f | %.2f | %.1f | %.3f | %.1f |" % (
        o["quads"], o["pairs_per_step"], d["inside_pairs_per_step"], d["k_pairs"], d["k_hull"], d["k_new_velocity"], d["kernels"],
        d["lqro_step_wall (host buffers)"], r["obstacle_creation"], r["half_plane_creation_without_hull"], r["hull"], r["new_velocity"], r["step"]))
