import sys; sys.path.insert(0,'/root/repo')
import ttcross_b200 as T
p = T.drivers.ising('c',10,256)
t=p.make(); t.set_partition(8)
g=t.dmrgg(32,p.accuracy,2)
for var in (1,0,2):
    r=t.superblock_probe(4, store=False, reps=5, variant=var)
    fl=r['count']*(5*p.d+3+2*32)
    print("variant",var,"ms",r['ms'],"Gevals/s",r['count']/r['ms']/1e6,"alg TFLOP/s",fl/r['ms']/1e9, r['argmax_a'], r['argmax_b'], r['a'], r['b'])
r=t.superblock_probe(4, store=True, reps=5, variant=0); print("stored ms", r['ms'], "GB/s", 8*r['count']/r['ms']/1e6)
print("fp64 peak DFMA TFLOP/s", T.fp64_peak(0, True), " DMUL+DADD TFLOP/s", T.fp64_peak(0, False))
