"""Tall-skinny Householder QR (ort0_d, reference lib/ort.f90:17-81 = LAPACK dgeqrf + dorgqr).

LAPACK is an un-vendored, unpinned dependency of the reference (`-llapack`, Makefile:18).  The oracle restates the
published unblocked algorithm (dgeqr2 / dlarfg / dlarf / dorg2r) and is PINNED here against numpy.linalg.qr, which calls
the same LAPACK routines the reference calls; the CUDA kernel is then compared with the oracle.  Summation order inside
LAPACK is build-specific, so the bar is rounding-level agreement (tolerances below), including LAPACK's sign convention."""
import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O

SHAPES = [(1, 1), (5, 1), (7, 3), (64, 8), (257, 16), (1040, 16), (2000, 32), (3, 5), (1, 4)]


def _mat(m, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((m, n)) * np.exp(rng.uniform(-3, 3, size=(1, n)))
    return np.asfortranarray(a)


def _check(a, q, r, tol):
    m, n = a.shape
    scale = np.linalg.norm(a)
    assert np.allclose(np.tril(r, -1), 0.0, atol=0)
    assert np.linalg.norm(q @ r - a) <= tol * scale
    assert np.linalg.norm(q.T @ q - np.eye(n)) <= tol * n


@pytest.mark.parametrize("m,n", SHAPES)
def test_oracle_qr_matches_lapack(m, n):
    a = _mat(m, n, 7 * m + n)
    q, r = O.qr_thin(a)
    if m < n:      # ort.f90:32-46: mat = [inp; 0], out = identity block
        assert np.array_equal(r[:m], a) and np.all(r[m:] == 0)
        assert np.array_equal(q, np.eye(m, n))
        return
    _check(a, q, r, 1e-13)
    ql, rl = np.linalg.qr(a, mode="reduced")               # LAPACK dgeqrf + dorgqr
    np.testing.assert_allclose(r, rl, rtol=0, atol=1e-12 * np.linalg.norm(a))
    np.testing.assert_allclose(q, ql, rtol=0, atol=1e-11)


def test_oracle_qr_rank_deficient_and_zero_column():
    a = _mat(50, 4, 3)
    a[:, 2] = 0.0                      # dlarfg with xnorm = 0 and alpha = 0: tau = 0, H = I
    q, r = O.qr_thin(a)
    assert np.linalg.norm(q @ r - a) <= 1e-13 * np.linalg.norm(a)
    rl = np.linalg.qr(a, mode="r")
    np.testing.assert_allclose(r, rl, rtol=0, atol=1e-12 * np.linalg.norm(a))


@pytest.mark.gpu
@pytest.mark.parametrize("m,n", SHAPES + [(8224, 32), (12336, 48), (32832, 64)])
def test_gpu_qr_matches_oracle(m, n):
    a = _mat(m, n, 11 * m + n)
    q, r, ms = T.qr_thin(a)
    qo, ro = O.qr_thin(a)
    if m < n:
        assert np.array_equal(q, qo) and np.array_equal(r, ro)
        return
    _check(a, q, r, 1e-13)
    np.testing.assert_allclose(r, ro, rtol=0, atol=1e-12 * np.linalg.norm(a))
    np.testing.assert_allclose(q, qo, rtol=0, atol=1e-11)


@pytest.mark.gpu
def test_gpu_qr_on_a_core_unfolding():
    """The caller this kernel exists for: the (r*n) x r unfolding of a TT core produced by the sweep (dtt_ort, tt.f90:130-198)."""
    p = T.drivers.ising("c", 6, 32)
    t = p.make()
    t.dmrgg(10, p.accuracy, 1)
    c = t.core(3)
    a = np.asfortranarray(c.reshape((c.shape[0] * c.shape[1], c.shape[2]), order="F"))
    q, r, _ = T.qr_thin(a)
    _check(a, q, r, 1e-13)
    qo, ro = O.qr_thin(a)
    np.testing.assert_allclose(r, ro, rtol=0, atol=1e-12 * np.linalg.norm(a))


def test_qr_fails_loudly_without_a_device(has_gpu):
    if has_gpu:
        pytest.skip("needs a box without a GPU")
    with pytest.raises(T.TTCrossError) as e:
        T.qr_thin(_mat(10, 3, 1))
    assert e.value.status == 4 and "no CPU fallback" in e.value.msg
