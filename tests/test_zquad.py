"""ztt_quad (reference lib/dmrgg.f90:1418-1523): quadrature of the train against complex rank-1 weights,
batched over weight sets the way test_crs_chf.f90:153-168 uses it (32 frequencies)."""
import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O


def _numpy_zquad(cores, w):
    v = np.ones((1,), dtype=complex)
    off = 0
    for c in cores:
        n = c.shape[1]
        v = v @ np.einsum("ijk,j->ik", c, w[off:off + n])
        off += n
    return v[0]


def test_oracle_zquad_matches_numpy():
    rng = np.random.default_rng(3)
    n, r = [7, 9, 8, 6], [1, 4, 6, 5, 1]
    cores = [rng.standard_normal((r[k], n[k], r[k + 1])) for k in range(4)]
    w = rng.standard_normal(sum(n)) + 1j * rng.standard_normal(sum(n))
    got, ref = O.quad_complex(cores, w), _numpy_zquad(cores, w)
    assert abs(got - ref) <= 1e-13 * abs(ref)
    wr = rng.standard_normal(sum(n))                      # real weights: the imaginary part vanishes exactly
    assert O.quad_complex(cores, wr).imag == 0.0


@pytest.mark.gpu
def test_gpu_zquad_batch_matches_oracle_and_dtt_quad():
    p = T.drivers.ising("c", 8, 32)
    t = p.make(); t.set_partition(2)
    t.dmrgg(12, p.accuracy, 2)
    cores = t.cores()
    nq = int(p.n[0])
    x, wq = p.par[:nq], p.quad[:nq]
    sets = []
    for k in range(32):                                    # the frequency loop of test_crs_chf.f90:153-166
        omega = k * np.pi / 300.0
        sets.append(np.tile(wq * np.exp(1j * omega * np.exp(x) / p.d), p.d))
    W = np.array(sets)
    got = t.quad_complex(W)
    want = np.array([O.quad_complex(cores, w) for w in W])
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
    assert got[0].imag == 0.0 and abs(got[0].real / t.quad() - 1) < 1e-13     # omega = 0: the real quadrature dtt_quad
