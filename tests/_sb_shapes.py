"""Superblock kernel (pivoting = -1 branch) at the shapes of the BASELINE configs + one full pivoting = -1 run of config B."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
import ttcross_b200 as T
pk = T.fp64_peak(0, False)
print("no-FMA FP64 ceiling TFLOP/s", pk)
for name, mk, R, piv, P in [("B", lambda: T.drivers.ising('c', 10, 256), 32, 2, 8), ("C", lambda: T.drivers.ising('d', 8, 256), 48, 2, 6),
                            ("D", lambda: T.drivers.ising('e', 6, 512), 64, 3, 4), ("E", lambda: T.drivers.mvn(64, 128), 32, 1, 63)]:
    p = mk(); t = p.make(); t.set_partition(P)
    g = t.dmrgg(R, p.accuracy, piv)
    bond = p.d // 2
    r0, r1, r2 = (int(g.ranks[b]) for b in (bond - 1, bond, bond + 1))
    d = p.d
    fl_eval = {"B": 5 * d + 3, "C": 3 * d * d + 8 * d + 4, "D": 3 * d * d + 4 * d + 1, "E": 3 * d * d + d + 10}[name]
    try:
        r = t.superblock_probe(bond, reps=2, variant=0)
        fl = r["count"] * (fl_eval + 2 * r1)
        line = f"{name}: shape {r0}x{int(p.n[0])}x{int(p.n[0])}x{r2} K={r1} elements {r['count']:.3e} tiled {r['ms']:.3f} ms = {r['count']/r['ms']/1e6:.1f} G evals/s, {fl/r['ms']/1e9:.2f} alg TFLOP/s ({100*fl/r['ms']/1e9/pk:.0f}% of ceiling)"
        if r["count"] < 3e8:
            q = t.superblock_probe(bond, reps=1, variant=1)
            line += f"; plain {q['ms']:.3f} ms; same argmax {q['argmax_b'] == r['argmax_b'] and q['b'] == r['b']}"
        print(line, flush=True)
        pkf = T.fp64_peak(0, True)
        for var, nm in ((2, "DFMA residual"), (3, "DMMA residual (fast mode)")):
            try:
                f = t.superblock_probe(bond, reps=2, variant=var)
                print(f"    {nm}: {f['ms']:.3f} ms, {fl/f['ms']/1e9:.2f} alg TFLOP/s = {100*fl/f['ms']/1e9/pkf:.0f}% of the DFMA peak {pkf:.1f}; "
                      f"same residual argmax {f['argmax_b'] == r['argmax_b']}, value rel diff {abs(f['b']/r['b']-1):.1e}", flush=True)
            except T.TTCrossError as e:
                print("   ", nm, "unavailable:", e.msg)
    except T.TTCrossError as e:
        print(name, "probe failed:", e)
p = T.drivers.ising('c', 10, 256)
for P in (8, 1):
    t = p.make(); t.set_partition(P)
    t0 = time.perf_counter(); g = t.dmrgg(32, p.accuracy, -1); w = time.perf_counter() - t0
    print(f"config B with pivoting=-1, P={P}: device {g.device_ms:.1f} ms wall {1e3*w:.1f} ms neval {g.neval:.4e} -> {g.neval/g.device_ms/1e6:.1f} G evals/s, val {g.vals[-1]!r} ranks {list(map(int,g.ranks))}", flush=True)
