"""Device time of one full TT-cross run of a BASELINE config on the GPUs of this job (1 process per GPU under torchrun).
usage: _cfg_time.py NAME   (A, B, C, D, E)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ttcross_b200 as T
name = sys.argv[1]
parts = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfgs = {"A": (lambda: T.drivers.ising("c", 6, 64), 16, 1, 4), "B": (lambda: T.drivers.ising("c", 10, 256), 32, 2, 8),
        "C": (lambda: T.drivers.ising("d", 8, 256), 48, 2, 6), "D": (lambda: T.drivers.ising("e", 6, 512), 64, 3, 4),
        "E": (lambda: T.drivers.mvn(64, 128), 32, 1, 63)}
mk, R, piv, P = cfgs[name]
if parts: P = parts
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
p = mk()
t = p.make(device=local); t.set_partition(P)
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local); dist.init_process_group("gloo")
    T.multi.attach(t, dist)
ms = []
for i in range(4):
    if dist: dist.barrier()
    t0 = time.perf_counter(); g = t.dmrgg(R, p.accuracy, piv); wall = time.perf_counter() - t0
    ms.append((g.device_ms, 1e3 * wall))
if dist:
    box = [None] * world; dist.all_gather_object(box, ms[-1]); dev = max(b[0] for b in box)
else:
    dev = ms[-1][0]
if rank == 0:
    print(f"config {name} P={P} gpus={world}: device ms {dev:.3f} (runs {[round(m[0],2) for m in ms]}) neval {g.neval} evals/s {g.neval/dev*1e3:.3e} sweeps {g.nsweeps} val {g.vals[-1]!r} ranks {list(map(int,g.ranks))[:6]}...")
t.close()
if dist: dist.destroy_process_group()
