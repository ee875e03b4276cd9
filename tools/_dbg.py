import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from gproshan_b200 import api, meshgen as mg
from oracle_lib import Oracle
o=Oracle()
m = mg.icosphere(400, noise_sigma=0.2 * mg.mean_edge_icosphere(400), seed=12345, dtype=np.float32)
src=[123456]
t,s,l=o.compute_toplesets(m,src); want,_,st=o.ptp_cpu(m,src,l,s); print(st)
with api.DeviceMesh(m,0) as dm:
    a,_,_=dm.geodesics(src); sa=dict(dm.last_stats)
    top,srt,lim=dm.compute_toplesets(src)
    b,_=dm.solve(src,lim,srt); sb=dict(dm.last_stats)
    c=dm.solve_batched(np.array(src,dtype=np.uint32))[0]; sc=dict(dm.last_stats)
for name,x,stt in (('fused',a,sa),('solve',b,sb),('batched',c,sc)):
    bad=np.nonzero(x!=want)[0]
    print(name,'nbad',bad.size,'iters',stt['iterations'],'upd',stt['vertex_updates'],'relax',stt['relaxations'], 'maxrel', (np.abs(x-want)/np.maximum(want,1e-30)).max())
    if bad.size: print('  levels of bad', np.unique(t[bad])[:20], 'first', bad[:5], x[bad[:5]], want[bad[:5]])
