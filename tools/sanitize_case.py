"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): 64 tight agents (hull branch taken), the
chunk-free host step, device kNN, two dynamics classes."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from lqr_obstacles_b200 import Lqro, swarm
from lqr_obstacles_b200.api import FLAG_DEVICE_KNN
g = np.load(os.path.join(R, "tests/golden/ref_vectors.npz"))
x, vp = swarm.lattice_swarm((4, 4, 4)); x[:, :3] *= 0.5
ctx = Lqro(n_agents=64, max_neighbors=16, horizon=20, flags=FLAG_DEVICE_KNN)
ctx.set_dynamics(g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"])
v = ctx.step(x, vp, None)
print("step ok", ctx.counters(), float(v.sum()))
ctx.set_dynamics_class(1, g["dyn_A"], g["dyn_B"], g["dyn_L"] * 0.9, g["dyn_E"] * 0.9)
ctx.set_agent_classes((np.arange(64) % 2).astype(np.int32))
v = ctx.step(x, vp, None)
print("classes ok", ctx.counters(), float(v.sum()))
pl, vt, deg = ctx.convex_hull(g["qh_points"])
print("hull ok", len(pl), deg)
ctx.close()
# a step large enough for the polling hull launch (concurrent consumer on its own stream next to the pair kernel, wide hull launch)
x, vp = swarm.lattice_swarm((16, 16, 8)); x[:, :3] *= 0.6
nbr = swarm.knn(x[:, :3], 64)
ctx = Lqro(n_agents=len(x), max_neighbors=64, horizon=20)
ctx.set_dynamics(g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"])
v = ctx.step(x, vp, nbr)
print("two-slice step ok", ctx.counters(), float(v.sum()))
# simulator-form update of the canonical 6 quads (golden state) and the batched Riccati solve
from lqr_obstacles_b200.api import SimModel
s = np.load(os.path.join(R, "tests/golden/sim_update.npz"))
ctx.sim_set_model(SimModel.from_array(s["model"]))
st = {k: s["s90_before_" + k].copy() for k in ("x", "Rot", "P", "xTrue", "RotTrue")}
vG, _ = ctx.sim_update(st["x"], st["Rot"], st["P"], st["xTrue"], st["RotTrue"], s["s90_newV"], s["s90_before_pGoal"], s["s90_normals"])
print("sim update ok", float(vG.sum()))
L, E, Lh, Eh = ctx.solve_gains(np.tile(g["dyn_A"], (8, 1, 1)), np.tile(g["dyn_B"], (8, 1, 1)), iters=20)
print("riccati ok", float(L.sum()))
ctx.close()
