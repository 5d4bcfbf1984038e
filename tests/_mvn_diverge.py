import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import ttcross_b200 as T
from parity_util import run_both
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
p = T.drivers.mvn(64, 128)
t, g, o = run_both(p, 32, 1, P=P)
n = min(len(g.pivlog), len(o.pivlog))
bad = [i for i in range(n) if not np.array_equal(g.pivlog[i], o.pivlog[i])]
print("records", len(g.pivlog), len(o.pivlog), "index mismatches", len(bad))
if bad:
    i = bad[0]
    print("first index mismatch at record", i)
    for k in range(max(0, i - 2), min(n, i + 3)):
        print(k, g.pivlog[k], repr(g.pivots[k]), "|", o.pivlog[k], repr(o.pivots[k]))
rel = np.abs(g.pivots[:n] - o.pivots[:n]) / np.maximum(np.abs(o.pivots[:n]), 1e-300)
print("max rel pivot diff before first mismatch", rel[: (bad[0] if bad else n)].max())
print("vals gpu", g.vals[-3:], "oracle", o.vals[-3:], "rel", abs(g.vals[-1] / o.vals[-1] - 1))
print("ranks equal", np.array_equal(g.ranks, o.ranks), "neval", g.neval, o.neval)
acc = (g.pivlog[:n, 7] == 1) & (o.pivlog[:n, 7] == 1)
lim = bad[0] if bad else n
for k in range(0, lim, max(1, lim // 24)):
    j = k
    while j < lim and not acc[j]: j += 1
    if j < lim: print("record", j, "sweep", g.pivlog[j][0], "rel diff", rel[j], "pivot", o.pivots[j], "amax", o.amaxs[min(g.pivlog[j][0], len(o.amaxs)-1)])
t2 = p.make(); t2.set_partition(P); t2.set_lottery_mode(4); g2 = t2.dmrgg(32, p.accuracy, 1)
print("cluster kernel vs split kernels identical:", np.array_equal(g.pivlog, g2.pivlog), np.array_equal(g.pivots, g2.pivots), np.array_equal(g.vals, g2.vals))
