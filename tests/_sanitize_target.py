"""Small end-to-end cases for compute-sanitizer (every kernel class: cluster visits, split kernels, superblock, exchange, quadrature, QR)."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import ttcross_b200 as T
from parity_util import run_both, first_pivot_mismatch
for (kind, idx, n, R, piv, P, mode) in [("c", 6, 16, 6, 2, 2, 0), ("d", 5, 12, 5, 1, 1, 0), ("c", 6, 16, 6, 2, 3, 4), ("c", 5, 12, 5, -1, 2, 0), ("e", 5, 12, 5, 0, 2, 0)]:
    p = T.drivers.ising(kind, idx, n)
    t = p.make(); t.set_partition(P); t.set_lottery_mode(mode)
    g = t.dmrgg(R, p.accuracy, piv)
    q = t.quad(); c = t.cores()
    t.superblock_probe(1 if P == 1 else 2, variant=0) if True else None
    print(kind, idx, n, R, piv, P, mode, "ok", g.neval, q)
p = T.drivers.mvn(4, 12); t = p.make(); g = t.dmrgg(5, p.accuracy, 1); print("mvn ok", g.neval)
a = np.asfortranarray(np.random.default_rng(0).standard_normal((300, 8)))
qq, rr, ms = T.qr_thin(a); print("qr ok", np.linalg.norm(qq @ rr - a))
