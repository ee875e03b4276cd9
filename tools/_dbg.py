import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from gproshan_b200 import api, meshgen as mg
from oracle_lib import Oracle
o=Oracle()
m=mg.grid(41).astype(np.float32); src=[20*41+20]
t,s,l=o.compute_toplesets(m,src); want,_,st=o.ptp_cpu(m,src,l,s)
with api.DeviceMesh(m,0) as dm:
    got,_,srt=dm.geodesics(src,want_sorted=True); print(dm.last_stats, st)
print('sorted equal', np.array_equal(srt,s[:l[-1]]))
bad=np.nonzero(got!=want)[0]; print('nbad',bad.size, 'of', got.size)
inv=np.empty(m.n_vertices,int); inv[s[:l[-1]]]=np.arange(l[-1])
lev=t
print('bad levels hist', np.bincount(lev[bad])[:45])
for v in bad[:10]: print(v, 'lvl',lev[v],'rank',inv[v], got[v], want[v])
