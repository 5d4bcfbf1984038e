import sys, time; sys.path.insert(0,'/root/repo')
from oracle import oracle as O
import ttcross_b200 as T, numpy as np
p = T.drivers.ising('c', 10, 256)
s = O.Setup(p.kind,p.d,p.n,p.par,p.aux,p.quad,p.accuracy,p.tru)
L = O.lib()
ref=None
for conc in (0,1):
    L.tto_set_rank_concurrency(conc)
    for nt in (1,2,4,8,16):
        L.tto_set_num_threads(nt)
        o=O.Oracle(s); t0=time.time(); r=o.run(maxrank=32,piv=2,P=8); dt=time.time()-t0
        if ref is None: ref=r
        same = np.array_equal(r.pivlog, ref.pivlog) and np.array_equal(r.vals, ref.vals) and np.array_equal(r.pivots, ref.pivots)
        print(f"conc={conc} threads={nt}: {dt:.3f}s  evals/s {r.neval/dt:.3e} identical={same}")
