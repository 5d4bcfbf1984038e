import sys, os; sys.path.insert(0,'/root/repo')
import ttcross_b200 as T
p = T.drivers.ising('e', 6, 512)
t = p.make(); t.set_partition(1)
try:
    g = t.dmrgg(64, p.accuracy, 3)
    print("ok", g.ranks, g.device_ms)
except Exception as e:
    print("FAIL", e)
