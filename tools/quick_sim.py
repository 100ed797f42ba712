"""Kernel times of the simulator-form update (k_sim_flow / k_sim_noise / k_sim_filter) for N agents (copies of the 6 golden
quads at step 90)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqr_obstacles_b200 import Lqro  # noqa: E402
from lqr_obstacles_b200.api import SimModel  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = np.load(os.path.join(ROOT, "tests", "golden", "sim_update.npz"))
reps = (N + 5) // 6
tile = lambda a: np.ascontiguousarray(np.tile(a, (reps, 1))[:N])
c = Lqro(n_agents=6, max_neighbors=5, horizon=30)
c.sim_set_model(SimModel.from_array(g["model"]))
ms = []
for it in range(4):
    st = [tile(g["s90_before_" + k]) for k in ("x", "Rot", "P", "xTrue", "RotTrue")]
    c.sim_update(*st, tile(g["s90_newV"]), tile(g["s90_before_pGoal"]), tile(g["s90_normals"]))
    ms.append(c.sim_last_ms(parts=True))
print("sim update N=%d ms (total, flow, noise, filter):" % N, [tuple(round(v, 3) for v in m) for m in ms], "agents/s %.3e" % (N / (min(m[0] for m in ms) * 1e-3)))
