"""Per-phase cycle profile of the hull kernel on the bench swarm (profiling build -DLQRO_HULL_PROF into
lqr_obstacles_b200/liblqro_hullprof.so; the product build carries no clock reads).  usage: python tools/hullprof.py [--build]"""
import os, subprocess, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'tests'))
prof_so = os.path.join(root, 'lqr_obstacles_b200', 'liblqro_hullprof.so')
if '--build' in sys.argv or not os.path.exists(prof_so):
    subprocess.check_call(['nvcc', '-ccbin', '/usr/bin/g++', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-fmad=false',
                           '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-DLQRO_HULL_PROF'] + ([os.environ['HULLPROF_EXTRA']] if os.environ.get('HULLPROF_EXTRA') else []) + ['-o', prof_so,
                           os.path.join(root, 'lqr_obstacles_b200', 'csrc', 'lqro_api.cu')])
    if '--build' in sys.argv: sys.exit(0)
os.environ['LQRO_LIB'] = prof_so
from lqr_obstacles_b200 import Lqro, swarm
g = np.load(os.path.join(root, 'tests', 'golden', 'ref_vectors.npz'))
dims = tuple(int(v) for v in os.environ.get('HULLPROF_DIMS', '64x32x32').split('x'))
x, vp = swarm.lattice_swarm(dims); nb = swarm.knn(x[:, :3], 64)
ctx = Lqro(n_agents=len(x), max_neighbors=64, horizon=20); ctx.set_dynamics(g['dyn_A'], g['dyn_B'], g['dyn_L'], g['dyn_E'])
for r in range(3):
    ctx.step(x, vp, nb if r == 0 else None)
print('kernel ms (pairs, hull, lp, total)', ctx.last_step_ms(), 'counters', ctx.counters())
c = ctx.hull_cycles()
names = ['init', 'partitionall', 'next+horizon', 'cone', 'planes+match', 'gather', 'walk', 'exhaustive+horizon test', 'append', 'ring+delete',
         'insertions', 'repartitioned', 'maxcycles', 'maxpts', 'maxfacets']
tot = c[:10].sum()
for n, v in zip(names, c): print('%-24s %12d %5.1f%%' % (n, v, 100 * v / tot if tot else 0))
res = ctx.pair_results(); st = res['status'] & 0xff
nr = res['nreach'][st == 2]; nf = res['nfacets'][st == 2]
print('inside hull sizes: n mean %.0f max %d; facets mean %.0f max %d' % (nr.mean(), nr.max(), nf.mean(), nf.max()))
print('cycles per insertion', tot / max(1, c[10]), 'points re-partitioned per insertion', c[11] / max(1, c[10]))
lg = ctx.hull_longest()
tl = lg[:10].sum()
print('LONGEST hull: %d points, %d facets, %d insertions, %d cycles (%.3f ms at 1.965 GHz), points re-partitioned %d' % (lg[12], lg[13], lg[10], tl, tl / 1.965e6, lg[11]))
for n, v in zip(names[:10], lg[:10]): print('  %-24s %12d %5.1f%%' % (n, v, 100 * v / tl if tl else 0))
