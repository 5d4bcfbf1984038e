"""Short single-GPU run of the bench workload for ncu captures (one warm-up run + one run).  argv[1]: lottery mode (0 default / 4 split kernels)."""
import sys; sys.path.insert(0, '/root/repo')
import ttcross_b200 as T
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pivoting = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = T.drivers.ising('c', 10, 256)
t = p.make(); t.set_partition(8); t.set_lottery_mode(mode)
for _ in range(2):
    g = t.dmrgg(32, p.accuracy, pivoting)
print("ok", g.device_ms, g.neval, g.launches)
if len(sys.argv) > 3:
    print(t.superblock_probe(4, store=False, reps=2))
