"""dtt_ort (reference lib/tt.f90:130-198), first row of SURVEY 8(f): left-to-right orthogonalisation of the train with the
tall-skinny QR kernel.  The oracle restatement is pinned against a step-by-step NumPy/LAPACK construction (numpy.linalg.qr
= dgeqrf + dorgqr, what the reference calls); the CUDA path is compared with the oracle and with the defining properties."""
import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O


def _rand_train(rng, n, r):
    return [rng.standard_normal((r[k], n[k], r[k + 1])) for k in range(len(n))]


def _full(cores):
    v = cores[0]
    for c in cores[1:]:
        v = np.tensordot(v, c, axes=([v.ndim - 1], [0]))
    return v


def _numpy_tt_ort(cores):
    """dtt_ort with LAPACK through NumPy, written from the reference text."""
    cores = [c.copy() for c in cores]
    d = len(cores)
    lognrm = 0.0
    for k in range(d - 1):
        r0, n, r1 = cores[k].shape
        a = cores[k].reshape((r0 * n, r1), order="F")
        q, mat = np.linalg.qr(a, mode="reduced")
        nrm = np.linalg.norm(mat)
        if nrm != 0:
            mat = mat / nrm
            lognrm += np.log(nrm)
        cores[k] = q.reshape((r0, n, r1), order="F")
        nxt = cores[k + 1]
        cores[k + 1] = (mat @ nxt.reshape((nxt.shape[0], -1), order="F")).reshape(nxt.shape, order="F")
    nrm = np.linalg.norm(cores[-1])
    if nrm != 0:
        cores[-1] = cores[-1] / nrm
        lognrm += np.log(nrm)
    s = np.exp(lognrm / d)
    return [s * c for c in cores]


def _check_properties(orig, new, tol):
    d = len(orig)
    a, b = _full(orig), _full(new)
    assert np.linalg.norm(a - b) <= tol * np.linalg.norm(a)                       # the same tensor
    norms = [np.linalg.norm(c) for c in new]
    for k in range(d - 1):                                                          # cores 1..d-1: orthonormal columns x common scale
        r0, n, r1 = new[k].shape
        u = new[k].reshape((r0 * n, r1), order="F")
        s2 = (u.T @ u)[0, 0]
        assert np.linalg.norm(u.T @ u - s2 * np.eye(r1)) <= tol * s2 * r1
    sc = norms[-1]                                                                  # every core carries the scale exp(lognrm / d)
    for k in range(d - 1):
        assert abs(norms[k] / (sc * np.sqrt(new[k].shape[2])) - 1) <= tol * 10


@pytest.mark.parametrize("n,r", [([5, 4, 6], [1, 3, 4, 1]), ([9] * 5, [1, 4, 7, 6, 3, 1]), ([33] * 4, [1, 8, 16, 8, 1])])
def test_oracle_tt_ort_matches_lapack_construction(n, r):
    rng = np.random.default_rng(sum(n) + sum(r))
    cores = _rand_train(rng, n, r)
    got = O.tt_ort(cores)
    ref = _numpy_tt_ort(cores)
    for g, e in zip(got, ref):
        np.testing.assert_allclose(g, e, rtol=0, atol=1e-11 * max(1.0, np.abs(e).max()))
    _check_properties(cores, got, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,index,n,R,piv,P", [("c", 6, 32, 10, 1, 1), ("c", 8, 16, 8, 2, 3), ("d", 5, 16, 8, 2, 1), ("c", 10, 64, 16, 2, 8)])
def test_gpu_tt_ort_matches_oracle(kind, index, n, R, piv, P):
    p = T.drivers.ising(kind, index, n)
    t = p.make(); t.set_partition(P)
    t.dmrgg(R, p.accuracy, piv)
    before = t.cores()
    q0 = t.quad()
    t.ort()
    after = t.cores()
    want = O.tt_ort(before)
    for g, e in zip(after, want):
        assert g.shape == e.shape
        np.testing.assert_allclose(g, e, rtol=0, atol=1e-10 * max(1e-300, np.abs(e).max()))
    _check_properties(before, after, 1e-11) if p.d <= 6 else None
    assert abs(t.quad() / q0 - 1) < 1e-11                                           # dtt_quad of the orthogonalised train: same integral
