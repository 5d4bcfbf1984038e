"""Config E (mvn 64 128 32, piv 1): how far the default-mode GPU run is from the oracle (final integral, ranks, neval),
and whether parity mode (deterministic exp on both sides) is bit-identical.  usage: _e_gap.py [P ...]"""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import ttcross_b200 as T
from parity_util import run_both, first_pivot_mismatch
p = T.drivers.mvn(64, 128)
for P in [int(a) for a in sys.argv[1:]] or [8, 63]:
    for mode in (0, 1):
        t0 = time.time()
        t, g, o = run_both(p, 32, 1, P=P, exp_mode=mode)
        mm = first_pivot_mismatch(g, o)
        q = t.quad()
        print(f"P={P} exp_mode={mode}: first mismatch {None if mm is None else (mm[0], int(g.pivlog[mm[0]][0]) if mm[0] < len(g.pivlog) else -1)} of {len(g.pivlog)} records; "
              f"ranks equal {np.array_equal(g.ranks, o.ranks)} (max |dr| {int(np.abs(g.ranks - o.ranks).max())}); neval {g.neval} vs {o.neval} (rel {abs(g.neval / o.neval - 1):.2e}); "
              f"val {g.vals[-1]!r} vs {o.vals[-1]!r} (rel {abs(g.vals[-1] / o.vals[-1] - 1):.2e}); quad {q!r} vs {o.quad_final!r} (rel {abs(q / o.quad_final - 1):.2e}); "
              f"|val-1| gpu {abs(g.vals[-1] - 1):.2e} oracle {abs(o.vals[-1] - 1):.2e}; vals bitwise {np.array_equal(g.vals, o.vals)}; "
              f"gpu ms {g.device_ms:.2f}; wall {time.time() - t0:.1f}s", flush=True)
        t.close()
