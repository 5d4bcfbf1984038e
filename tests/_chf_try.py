import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ttcross_b200 as T
REF = [(1.0000000000000004, 0.0), (0.477573486783962, 0.8149466884621303), (-0.36837299289083225, 0.7151922711168992), (-0.6257714195536066, 0.09178553732163314), (-0.34684735452966187, -0.3172332243460625), (-0.0017955158585152564, -0.33646866663071234), (0.162377557153463, -0.17003135351832718), (0.16111757391061166, -0.014897840710532378)]
for d in (2, 3, 4, 6, 8):
    p = T.drivers.mvn(d, 64)
    t = p.make(use_quad=False, use_tru=False)
    g = t.dmrgg(20, p.accuracy, 1)
    nq = int(p.n[0]); x, w = p.par[:nq], p.par[nq:2*nq]
    W = np.array([np.tile(w * np.exp(1j * (k * np.pi / 300.0) * np.exp(x) / d), d) for k in range(8)])
    got = t.quad_complex(W)
    ref = np.array([complex(np.float32(a), np.float32(b)) for a, b in REF])
    print(d, "ranks", list(g.ranks), "max |got-ref|", np.abs(got - ref).max(), "got[1]", got[1], "ref[1]", ref[1])
