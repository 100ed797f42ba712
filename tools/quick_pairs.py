"""Kernel-only timing of the avoidance step on the bench swarm (resident neighbour table after the first step).
usage: python tools/quick_pairs.py [dims K]   env: LQRO_LIB, LQRO_HULL_OVERLAP_SMS"""
import sys, os, numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from lqr_obstacles_b200 import Lqro, swarm
g = np.load(os.path.join(root, 'tests', 'golden', 'ref_vectors.npz'))
dims = tuple(int(v) for v in sys.argv[1].split('x')) if len(sys.argv) > 1 else (64, 32, 32)
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
x, vp = swarm.lattice_swarm(dims); nb = swarm.knn(x[:, :3], K)
ctx = Lqro(n_agents=len(x), max_neighbors=K, horizon=20); ctx.set_dynamics(g['dyn_A'], g['dyn_B'], g['dyn_L'], g['dyn_E'])
ms = []
for r in range(8):
    nv = ctx.step(x, vp, nb if r == 0 else None)
    if r >= 2: ms.append(ctx.last_step_ms())
med = {k: float(np.median([m[k] for m in ms])) for k in ms[0]}
print("overlap_sms", os.environ.get("LQRO_HULL_OVERLAP_SMS"), "first_slots", os.environ.get("LQRO_FIRST_SLOTS"), ctx.last_step_detail_ms(), 'agents', len(x), 'K', K, 'median ms', {k: round(v, 4) for k, v in med.items()},
      'checksum', float(nv.sum()), ctx.counters())
