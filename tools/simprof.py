"""Per-phase cycle profile of the simulator-form update kernel (profiling build -DLQRO_SIM_PROF into
lqr_obstacles_b200/liblqro_simprof.so).  usage: python tools/simprof.py [--build] [N]"""
import os, subprocess, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
prof_so = os.path.join(root, 'lqr_obstacles_b200', 'liblqro_simprof.so')
if '--build' in sys.argv or not os.path.exists(prof_so):
    subprocess.check_call(['nvcc', '-ccbin', '/usr/bin/g++', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-fmad=false',
                           '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-DLQRO_SIM_PROF', '-o', prof_so,
                           os.path.join(root, 'lqr_obstacles_b200', 'csrc', 'lqro_api.cu')])
    if '--build' in sys.argv: sys.exit(0)
os.environ['LQRO_LIB'] = prof_so
from lqr_obstacles_b200 import Lqro
from lqr_obstacles_b200.api import SimModel
args = [a for a in sys.argv[1:] if not a.startswith('--')]
N = int(args[0]) if args else 4736
g = np.load(os.path.join(root, 'tests', 'golden', 'sim_update.npz'))
reps = (N + 5) // 6
tile = lambda a: np.ascontiguousarray(np.tile(a, (reps, 1))[:N])
c = Lqro(n_agents=6, max_neighbors=5, horizon=30)
c.sim_set_model(SimModel.from_array(g['model']))
for it in range(2):
    c.sim_cycles(reset=True)
    st = [tile(g['s90_before_' + k]) for k in ('x', 'Rot', 'P', 'xTrue', 'RotTrue')]
    c.sim_update(*st, tile(g['s90_newV']), tile(g['s90_before_pGoal']), tile(g['s90_normals']))
cyc = c.sim_cycles()
names = ['findU', 'Jacobian (f)', 'exp: powers+Pade', 'exp: solve', 'exp: squarings', 'noise cov + Simpson', 'Jacobi16 + sample',
         'propagate tail', 'KF1 tail', 'observation sample', 'KF2', 'findVGoal']
tot = cyc[:12].sum()
print('N', N, 'kernel ms (total, flow, noise, filter)', c.sim_last_ms(parts=True), 'cycles per agent', tot / N)
for n, v in zip(names, cyc): print('%-22s %12.0f cyc/agent %5.1f%%' % (n, v / N, 100 * v / tot))
