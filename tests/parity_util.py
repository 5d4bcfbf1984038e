"""Shared helpers of the parity tests: run the CUDA path (through the C-ABI) and the CPU oracle on the same
problem, seed and partition, and compare pivot tape, ranks, evaluation counts, per-sweep values and cores."""
from __future__ import annotations

import numpy as np

import ttcross_b200 as T
from oracle import oracle as O


def to_oracle_setup(p: "T.drivers.Problem") -> "O.Setup":
    return O.Setup(p.kind, p.d, p.n, p.par, p.aux, p.quad, p.accuracy, p.tru, p.label)


def run_both(p, maxrank, piv, P=1, own=None, seed=1, accuracy=None, use_quad=True, use_tru=True, exp_mode=0):
    acc = p.accuracy if accuracy is None else accuracy
    t = p.make(use_quad=use_quad, use_tru=use_tru)
    t.set_partition(P, own)
    t.set_seed(seed)
    t.set_exp_mode(exp_mode)          # 1: both sides evaluate exp through include/ttc_detexp.h (parity mode)
    g = t.dmrgg(maxrank, acc, piv)
    orc = O.Oracle(to_oracle_setup(p))
    orc.set_exp_mode(exp_mode)
    o = orc.run(maxrank=maxrank, piv=piv, P=P, own=own, accuracy=acc, use_quad=use_quad,
                                         use_tru=use_tru, seed=seed)
    return t, g, o


def first_pivot_mismatch(g, o):
    n = min(len(g.pivlog), len(o.pivlog))
    for i in range(n):
        if not np.array_equal(g.pivlog[i], o.pivlog[i]) or g.pivots[i] != o.pivots[i]:
            return i, g.pivlog[i], g.pivots[i], o.pivlog[i], o.pivots[i]
    if len(g.pivlog) != len(o.pivlog):
        return n, None, None, None, None
    return None


def assert_parity(t, g, o, exact=True, rtol=0.0):
    """exact: bit-identical pivots/values/cores (the +-*/ integrands).  Otherwise pivot indices identical and
    values within rtol (exp-based integrands; CUDA exp and glibc exp differ in the last ulp)."""
    assert o.status == 0
    mm = first_pivot_mismatch(g, o)
    if exact:
        assert mm is None, f"pivot tape differs at record {mm}"
    else:
        assert len(g.pivlog) == len(o.pivlog)
        assert np.array_equal(g.pivlog, o.pivlog), f"pivot indices differ; first mismatch {mm}"
        acc = g.pivlog[:, 7] == 1
        np.testing.assert_allclose(g.pivots[acc], o.pivots[acc], rtol=rtol, atol=0)
        # rejected candidates are residuals at rounding-noise level: only their magnitude is meaningful
        if (~acc).any():
            assert np.abs(g.pivots[~acc]).max() <= 1e-5 * max(np.abs(o.amaxs).max(), 1e-300)
    assert g.nsweeps == o.nsweeps
    assert np.array_equal(g.ranks, o.ranks)
    assert g.neval == o.neval
    assert np.array_equal(g.nevals, o.nevals)
    if exact:
        assert np.array_equal(g.vals, o.vals), f"per-sweep values differ: {g.vals - o.vals}"
        assert np.array_equal(g.amaxs, o.amaxs)
        assert np.array_equal(g.pivotmaxs, o.pivotmaxs)
    else:
        np.testing.assert_allclose(g.vals, o.vals, rtol=rtol)
    q = t.quad()
    if exact:
        assert q == o.quad_final, f"final quadrature {q!r} vs {o.quad_final!r}"
    else:
        np.testing.assert_allclose(q, o.quad_final, rtol=rtol)
    for k in range(1, t.d + 1):
        c = t.core(k)
        assert c.shape == o.cores[k - 1].shape
        if exact:
            assert np.array_equal(c, o.cores[k - 1]), f"core {k} differs, max abs diff {np.abs(c - o.cores[k-1]).max()}"
        else:
            scale = np.abs(o.cores[k - 1]).max()
            np.testing.assert_allclose(c, o.cores[k - 1], rtol=0, atol=1e-9 * scale)
    return q
