import sys; sys.path.insert(0,'/root/repo')
import numpy as np, collections
import ttcross_b200 as T
import os
cfg = os.environ.get('TTC_TL_CFG', 'B')
if cfg == 'E':
    p = T.drivers.mvn(64, 128); R, piv, P0 = 32, 1, 63
else:
    p = T.drivers.ising('c',10,256); R, piv, P0 = 32, 2, 8
t=p.make(); t.set_partition(int(sys.argv[1]) if len(sys.argv)>1 else P0)
for i in range(3): g=t.dmrgg(R,p.accuracy,piv)
print("device ms", g.device_ms, "sweeps", g.nsweeps)
t.set_timeline(True)
g=t.dmrgg(R,p.accuracy,piv)
tl=t.timeline()
print("timeline device ms", g.device_ms, "stamps", len(tl))
# intervals: time from this stamp to the next stamp, attributed to this kernel
agg=collections.defaultdict(list)
for (a,ta),(b,tb) in zip(tl[:-1],tl[1:]): agg[a].append((tb-ta)/1e3)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k:24s} n={len(v):5d} total={sum(v):9.1f} us avg={np.mean(v):7.2f} med={np.median(v):7.2f} share={100*sum(v)/tot:5.1f}%")
# one late sweep in detail
names=[a for a,_ in tl]
idx=[i for i,a in enumerate(names) if a=='k_visits']
if len(idx) < 3: idx=[i for i,a in enumerate(names) if a=='v:staged']     # persistent kernel: one launch, sweeps start at 'v:staged'
i0,i1=idx[-3],idx[-2]
t0=tl[i0][1]
ck=t.timeline_clocks
prev=None
for a,ta in tl[i0:i1+1]:
    extra=''
    if prev is not None and a.startswith(('v:','f:','q:','s:','l:')) and prev[0].startswith(('v:','f:','q:','s:','l:','k_visits','k_quad_inc')):
        dc=ck[ta]-ck[prev[1]]; dt=ta-prev[1]
        if dt>0 and 0<dc<10**7: extra=f"  dcycles={dc} -> {dc/dt*1e3:.0f} MHz"
    print(f"  {(ta-t0)/1e3:8.2f} us  {a}{extra}")
    prev=(a,ta)
