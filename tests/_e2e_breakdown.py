import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
import ttcross_b200 as T
p = T.drivers.ising('c',10,256)
def once(tag):
    t0=time.perf_counter(); t=p.make(); t.set_partition(8); t1=time.perf_counter()
    g=t.dmrgg(32,p.accuracy,2); t2=time.perf_counter()
    c=t.cores(); t3=time.perf_counter(); q=t.quad(); t4=time.perf_counter(); t.close(); t5=time.perf_counter()
    print(f"{tag}: make {1e3*(t1-t0):.2f} dmrgg {1e3*(t2-t1):.2f} (device {g.device_ms:.2f}) cores {1e3*(t3-t2):.2f} quad {1e3*(t4-t3):.2f} close {1e3*(t5-t4):.2f} total {1e3*(t5-t0):.2f} ms launches {g.launches}")
    return g
for i in range(4): g=once(f"run{i}")
t=p.make(); t.set_partition(8)
for i in range(3):
    t0=time.perf_counter(); g=t.dmrgg(32,p.accuracy,2); print(f"resident {i}: wall {1e3*(time.perf_counter()-t0):.2f} device {g.device_ms:.2f}")
t.set_partition(1)
for i in range(2):
    t0=time.perf_counter(); g=t.dmrgg(32,p.accuracy,2); print(f"P=1 resident {i}: wall {1e3*(time.perf_counter()-t0):.2f} device {g.device_ms:.2f} neval {g.neval} evals/s {g.neval/g.device_ms*1e3:.3e}")
