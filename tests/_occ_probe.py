import os, sys
sys.path.insert(0, '/root/repo')
os.environ['TTC_TRACE'] = '1'
import ttcross_b200 as T
p = T.drivers.ising('c', 10, 256)
for cs, th in ((16, 256), (16, 128), (8, 256), (16, 224), (16, 192)):
    os.environ['TTC_CLUSTER_SIZE'] = str(cs); os.environ['TTC_CLUSTER_THREADS'] = str(th)
    t = p.make(); t.set_partition(8)
    g = t.dmrgg(32, p.accuracy, 2)
    g = t.dmrgg(32, p.accuracy, 2)
    print(f"cluster {cs} x {th}: device ms {g.device_ms:.3f} launches {g.launches}", flush=True)
    t.close()
