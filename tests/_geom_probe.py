"""Device time of config B for several cluster geometries / register caps of the persistent kernel (one subprocess each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys
sys.path.insert(0, %r)
import ttcross_b200 as T
cfg = os.environ.get("CFG", "B")
mk = {"A": (lambda: T.drivers.ising("c", 6, 64), 16, 1, 4), "B": (lambda: T.drivers.ising("c", 10, 256), 32, 2, 8),
      "C": (lambda: T.drivers.ising("d", 8, 256), 48, 2, 6), "D": (lambda: T.drivers.ising("e", 6, 512), 64, 3, 4)}[cfg]
p = mk[0](); t = p.make(); t.set_partition(mk[3])
for i in range(4): g = t.dmrgg(mk[1], p.accuracy, mk[2])
print(f"   config {cfg}: device ms {g.device_ms:.3f} launches {g.launches} val {g.vals[-1]!r}")
''' % ROOT
variants = [("default", None, 16, 128), ("minb2", "build/lib_minb2.so", 16, 256), ("t192", "build/lib_t192.so", 16, 192),
            ("t352", "build/lib_t352.so", 12, 352), ("t352", "build/lib_t352.so", 14, 320), ("t352", "build/lib_t352.so", 14, 352),
            ("t352", "build/lib_t352.so", 8, 352), ("default", None, 8, 256), ("minb2", "build/lib_minb2.so", 8, 256)]
for name, lib, cs, th in variants:
    env = dict(os.environ, TTC_TRACE="1", TTC_CLUSTER_SIZE=str(cs), TTC_SWEEP_THREADS=str(th), TTC_CLUSTER_THREADS=str(min(th, 256)))
    if lib: env["TTC_LIB_PATH"] = os.path.join(ROOT, lib)
    for cfg in sys.argv[1:] or ["B"]:
        env["CFG"] = cfg
        r = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=300)
        tr = [l for l in r.stderr.splitlines() if "persistent kernel" in l][-1:] or [r.stderr[-300:]]
        print(f"{name} {cs}x{th}: {tr[0].replace('[ttc trace] ', '')}\n{r.stdout.rstrip()}", flush=True)
