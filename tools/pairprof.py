"""Per-phase cycle profile of the pair kernel (needs the profiling build: tools/pairprof.py builds it with
-DLQRO_PAIR_PROF into gpurun_out-independent lqr_obstacles_b200/liblqro_prof.so and loads that)."""
import os, subprocess, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
prof_so = os.path.join(root, 'lqr_obstacles_b200', 'liblqro_prof.so')
if '--build' in sys.argv or not os.path.exists(prof_so):
    subprocess.check_call(['nvcc', '-ccbin', '/usr/bin/g++', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-fmad=false',
                           '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-DLQRO_PAIR_PROF', '-o', prof_so,
                           os.path.join(root, 'lqr_obstacles_b200', 'csrc', 'lqro_api.cu')])
    if '--build' in sys.argv: sys.exit(0)
os.environ['LQRO_LIB'] = prof_so
from lqr_obstacles_b200 import Lqro, swarm
g = np.load(os.path.join(root, 'tests/golden/ref_vectors.npz'))
x, vp = swarm.lattice_swarm((64, 32, 32)); nb = swarm.knn(x[:, :3], 64)
ctx = Lqro(n_agents=65536, max_neighbors=64, horizon=20); ctx.set_dynamics(g['dyn_A'], g['dyn_B'], g['dyn_L'], g['dyn_E'])
for r in range(3): ctx.step(x, vp, nb)
print(ctx.last_step_ms(), ctx.counters())
c = ctx.hull_cycles()
names = ['batch+stage1', 'setup+translate', 'scan', 'gjk', 'output', '-']
tot = c[:6].sum()
for n, v in zip(names, c): print('%-20s %14d %5.1f%%' % (n, v, 100 * v / tot if tot else 0))
gn = ['step size 2', 'step size 3', 'step size 4', 'witness + d^2', 'support scan', 'arg-max, G-test, new point', 'steps of size 3', 'steps of size 4']
gt = c[8:14].sum()
for n, v in zip(gn, c[8:16]): print('  gjk: %-28s %14d %s' % (n, v, ('%5.1f%%' % (100 * v / gt)) if n[:5] != 'steps' and gt else ''))
print('stage-2 pairs', c[6], 'reachable points', c[7], 'cycles per stage-2 pair', tot / max(1, c[6]), 'points per pair', c[7] / max(1, c[6]))
