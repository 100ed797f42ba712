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
