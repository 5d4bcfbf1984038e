"""Device time of the thin QR (ttc_qr_thin) at the unfolding shapes of the BASELINE configs: TSQR vs the grid-wide panel kernel."""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import ttcross_b200 as T
rng = np.random.default_rng(1)
for m, n in [(8224, 32), (12336, 48), (32832, 64), (2000, 32)]:
    a = np.asfortranarray(rng.standard_normal((m, n)) * np.exp(rng.uniform(-3, 3, size=(1, n))))
    out = []
    for env in ("", "1"):
        if env: os.environ["TTC_NO_TSQR"] = env
        else: os.environ.pop("TTC_NO_TSQR", None)
        q, r, ms = T.qr_thin(a, reps=5)
        ql, rl = np.linalg.qr(a)
        out.append((ms, np.abs(r - rl).max() / np.linalg.norm(a), np.abs(q - ql).max(), np.linalg.norm(q.T @ q - np.eye(n))))
    print(f"{m} x {n}: TSQR {out[0][0]:.3f} ms (|R-R_lapack|/|A| {out[0][1]:.1e}, |Q-Q_lapack| {out[0][2]:.1e}, |QtQ-I| {out[0][3]:.1e});  panel kernel {out[1][0]:.3f} ms ({out[1][1]:.1e}, {out[1][2]:.1e}, {out[1][3]:.1e})", flush=True)
