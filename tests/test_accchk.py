"""dtt_ijk (reference lib/tt.f90:630-652) and dtt_accchk (lib/dmrgg.f90:1081-1166) on the device — SURVEY 8(f) rank 4 (checker half).
Checked against NumPy contractions of the cores the library returns and the CPU oracle's integrand."""
import ctypes as C

import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _np_value(cores, idx):
    v = np.ones((1, 1))
    for c, i in zip(cores, idx):
        v = v @ c[:, i - 1, :]
    return float(v[0, 0])


@pytest.mark.parametrize("kind,index,n,R,P", [("c", 6, 32, 10, 1), ("d", 5, 16, 8, 2), ("c", 10, 64, 16, 8)])
def test_values_match_numpy(kind, index, n, R, P):
    p = T.drivers.ising(kind, index, n)
    t = p.make(); t.set_partition(P)
    t.dmrgg(R, p.accuracy, 2)
    cores = t.cores()
    rng = np.random.default_rng(5)
    ind = rng.integers(1, int(p.n[0]) + 1, size=(500, p.d)).astype(np.int32)
    got = t.values(ind)
    want = np.array([_np_value(cores, row) for row in ind])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14 * np.abs(want).max())
    with pytest.raises(T.TTCrossError, match="out of range"):
        t.values(np.full((1, p.d), int(p.n[0]) + 1))


def test_accchk_matches_host_recomputation():
    p = T.drivers.ising("c", 6, 32)
    t = p.make()
    g = t.dmrgg(12, p.accuracy, 2)
    nlot, seed = 4000, 11
    r = t.accchk(nlot, seed)
    L = T.load_library()
    lib = O.lib()
    h = lib.tto_create(p.kind, p.d, p.n.ctypes.data_as(C.POINTER(C.c_int)), p.par.ctypes.data_as(C.POINTER(C.c_double)), p.par.size, None, 0)
    try:
        ind = np.zeros((nlot, p.d), dtype=np.int32)
        for x in range(nlot):
            for q in range(p.d):
                ind[x, q] = min(int(L.ttc_stream_uniform(seed, 0x7fffffff, x * p.d + q) * int(p.n[q])) + 1, int(p.n[q]))   # irnd, rnd.f90:84-90
        a = np.array([lib.tto_integrand(h, row.ctypes.data_as(C.POINTER(C.c_int))) for row in ind])
    finally:
        lib.tto_destroy(h)
    b = t.values(ind)
    e = np.abs(a - b)
    assert abs(r["einf"] - e.max()) <= 1e-12 * max(e.max(), 1e-300)
    assert list(r["pivot"]) == list(ind[int(np.argmax(e))])
    assert abs(r["efro"] / np.sqrt((e ** 2).sum()) - 1) < 1e-10
    assert r["ainf"] == a.max() and abs(r["afro"] / np.sqrt((a ** 2).sum()) - 1) < 1e-12
    assert r["efro"] / r["afro"] < 1e-6           # the rank-12 cross of C_6 on 33 nodes is accurate to ~1e-8 in the Frobenius sense
