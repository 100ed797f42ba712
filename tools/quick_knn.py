import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lqr_obstacles_b200 import Lqro, swarm
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/ref_vectors.npz'))
x, vp = swarm.lattice_swarm((64, 32, 32))
ctx = Lqro(n_agents=65536, max_neighbors=64, horizon=20); ctx.set_dynamics(g['dyn_A'], g['dyn_B'], g['dyn_L'], g['dyn_E'])
nb = ctx.knn(x)
for rep in range(2):
    ctx.synchronize(); t0 = time.perf_counter()
    for r in range(20): ctx.knn_device()
    ctx.synchronize(); print('knn_device ms', (time.perf_counter() - t0) / 20 * 1e3)
