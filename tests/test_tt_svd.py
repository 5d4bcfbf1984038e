"""dtt_svd (reference lib/tt.f90:307-368 with d_svd / chop of lib/mat.f90:340-385, 433-455): TT rounding — second half of the first
row of SURVEY 8(f).  The oracle (one-sided Jacobi in place of the unpinned LAPACK dgesvd) is pinned against a NumPy/LAPACK
construction written from the reference text; the CUDA path (QR of the transposed unfolding + Jacobi on the small factor) is
compared with the oracle through ranks, the rounded tensor and the integral."""
import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O
from test_tt_ort import _numpy_tt_ort, _full, _rand_train


def _chop(s, tol, rmax):
    r, er2 = len(s), 0.0
    if rmax and rmax < r:
        er2 = float(s[rmax:] @ s[rmax:]); r = rmax
    if tol is not None and tol >= 0:
        bound = tol * tol * float(s @ s)
        er = er2 + s[r - 1] ** 2
        while er < bound and r > 1:
            er2 = er; r -= 1; er += s[r - 1] ** 2
    return r


def _numpy_tt_svd(cores, tol, rmax):
    cores = _numpy_tt_ort(cores)
    d = len(cores)
    lognrm = 0.0
    for k in range(d - 1, 0, -1):
        r0, n, r1 = cores[k].shape
        u, s, vt = np.linalg.svd(cores[k].reshape((r0, n * r1), order="F"), full_matrices=False)
        rr = _chop(s, tol, rmax)
        s = s[:rr]
        nrm = np.linalg.norm(s)
        if nrm != 0:
            s = s / nrm; lognrm += np.log(nrm)
        us = u[:, :rr] * s
        p = cores[k - 1]
        cores[k - 1] = (p.reshape((-1, r0), order="F") @ us).reshape((p.shape[0], p.shape[1], rr), order="F")
        cores[k] = vt[:rr].reshape((rr, n, r1), order="F")
    nrm = np.linalg.norm(cores[0])
    if nrm != 0:
        cores[0] = cores[0] / nrm; lognrm += np.log(nrm)
    sc = np.exp(lognrm / d)
    return [sc * c for c in cores]


@pytest.mark.parametrize("tol,rmax", [(1e-10, 0), (1e-3, 0), (-1.0, 3), (1e-12, 4)])
def test_oracle_tt_svd_matches_lapack_construction(tol, rmax):
    rng = np.random.default_rng(7)
    n, r = [6, 5, 7, 6], [1, 5, 9, 6, 1]
    cores = _rand_train(rng, n, r)
    cores[1] = cores[1] * np.exp(-2.0 * np.arange(r[2]))[None, None, :]      # decaying singular values across bond 2
    got = O.tt_svd(cores, tol, rmax)
    ref = _numpy_tt_svd(cores, tol if tol >= 0 else None, rmax)
    assert [c.shape for c in got] == [c.shape for c in ref]
    a, b = _full(got), _full(ref)
    assert np.linalg.norm(a - b) <= 1e-11 * np.linalg.norm(b)
    if tol >= 0 and not rmax:
        assert np.linalg.norm(a - _full(cores)) <= 4 * tol * np.linalg.norm(_full(cores)) + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kind,index,n,R,tol,rmax", [("c", 6, 32, 12, 1e-6, 0), ("c", 6, 32, 12, -1.0, 5), ("d", 5, 16, 8, 1e-8, 0), ("c", 8, 24, 10, 1e-4, 6)])
def test_gpu_tt_svd_matches_oracle(kind, index, n, R, tol, rmax):
    p = T.drivers.ising(kind, index, n)
    t = p.make()
    t.dmrgg(R, p.accuracy, 2)
    before = t.cores()
    q0 = t.quad()
    want = O.tt_svd(before, tol, rmax)
    t.svd(tol, rmax)
    after = t.cores()
    assert [c.shape for c in after] == [c.shape for c in want], ([c.shape for c in after], [c.shape for c in want])
    rng = np.random.default_rng(1)
    ind = rng.integers(1, int(p.n[0]) + 1, size=(400, p.d)).astype(np.int32)
    def val(cores, row):
        v = np.ones((1, 1))
        for c, i in zip(cores, row):
            v = v @ c[:, i - 1, :]
        return v[0, 0]
    a = t.values(ind)
    b = np.array([val(want, row) for row in ind])
    o = np.array([val(before, row) for row in ind])
    scale = np.abs(o).max()
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-9 * scale)                       # the same rounded tensor
    eff = tol if tol >= 0 else 1.0
    assert abs(t.quad() / q0 - 1) < max(50 * eff, 1e-9) if not rmax else True           # the integral moves by O(tol)
    assert max(t.ranks) <= max(c.shape[2] for c in before)
