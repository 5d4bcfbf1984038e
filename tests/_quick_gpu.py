import sys, traceback; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import ttcross_b200 as T
from parity_util import run_both, first_pivot_mismatch
for (kind,idx,n,R,piv,P) in [("c",4,8,6,1,1),("c",6,64,16,1,1),("d",5,16,8,2,1),("c",5,16,8,0,1),("c",5,12,6,-1,1),("c",6,16,8,1,2),("c",10,32,10,2,8)]:
    try:
        p = T.drivers.ising(kind, idx, n)
        t,g,o = run_both(p,R,piv,P=P)
        print(kind,idx,n,R,piv,P,'mismatch:',first_pivot_mismatch(g,o),'ranks',g.ranks,o.ranks,'neval',g.neval,o.neval)
        print('  vals equal', np.array_equal(g.vals,o.vals), g.vals[-1], o.vals[-1], 'quad', t.quad(), o.quad_final)
        for k in range(1,t.d+1):
            c=t.core(k)
            if c.shape!=o.cores[k-1].shape or not np.array_equal(c,o.cores[k-1]): print('  core',k,'differs', c.shape, o.cores[k-1].shape, np.abs(c-o.cores[k-1]).max() if c.shape==o.cores[k-1].shape else '')
        print('  gpu ms', g.device_ms, 'launches', g.launches, 'oracle s', o.seconds)
    except Exception as e:
        traceback.print_exc()
