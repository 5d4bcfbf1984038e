"""ncu target: ONE thin QR at the config-D unfolding shape (32832 x 64) -- for the per-kernel launch list of the TSQR."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import ttcross_b200 as T
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32832, 64)
rng = np.random.default_rng(1)
a = np.asfortranarray(rng.standard_normal((m, n)) * np.exp(rng.uniform(-3, 3, size=(1, n))))
q, r, ms = T.qr_thin(a, reps=2)
print(m, n, ms)
