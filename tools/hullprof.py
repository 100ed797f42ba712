import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from lqr_obstacles_b200 import Lqro, swarm
g=np.load('/root/repo/tests/golden/ref_vectors.npz')
x,vp=swarm.lattice_swarm((64,32,32)); nb=swarm.knn(x[:,:3],64)
ctx=Lqro(n_agents=65536,max_neighbors=64,horizon=20); ctx.set_dynamics(g['dyn_A'],g['dyn_B'],g['dyn_L'],g['dyn_E'])
for r in range(3):
    ctx.step(x,vp,nb if r==0 else None)
print(ctx.last_step_ms(), ctx.counters())
c=ctx.hull_cycles()
names=['init','partall','next+horizon+cone','planes+match','list','walk','redo','append','delete','append-old','insertions','repartitioned','maxcycles','maxpts','maxfacets']
tot=c[:10].sum()
for n,v in zip(names,c): print('%-20s %12d %5.1f%%'%(n,v,100*v/tot if tot else 0))
res=ctx.pair_results(); st=res['status']&0xff
nr=res['nreach'][st==2]; nf=res['nfacets'][st==2]
print('inside hull sizes: n mean %.0f max %d; facets mean %.0f max %d'%(nr.mean(),nr.max(),nf.mean(),nf.max()))
print('cycles per insertion', tot/max(1,c[10]), 'per repartitioned point', (c[4]+c[5]+c[6]+c[7])/max(1,c[11]))
