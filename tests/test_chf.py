"""The reference's characteristic-function driver (test_crs_chf.f90): MVN cross without quad/tru (:122-123), then 32 complex
quadratures ztt_quad of the same train (:153-168).  Its own known-answer table get_reference_val (:232-271) is the pin:
the values belong to DIM = 4 (found by running the pipeline for several DIM; the driver's default DIM = 6 does not
reproduce them) and are single-precision literals whose tail digits look sampled, so the tolerance is absolute 2e-4
(the oracle lands within 8e-5, tests/golden/reference/make_chf_table.py extracted the table)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TOL = 2e-4


def _table():
    vals = json.load(open(os.path.join(HERE, "golden", "reference", "chf_table.json")))["values"]
    return np.array([complex(np.float32(a), np.float32(b)) for a, b in vals])     # default-kind cmplx() rounds to single


def _weights(x, w, d):
    return np.array([np.tile(w * np.exp(1j * (k * np.pi / 300.0) * np.exp(x) / d), d) for k in range(32)])


def test_cos_approximate_recovers_a_density_from_the_reference_table():
    """test_crs_pdf.f90:171-183: the COS series of the 32 table values on [0, 300] is the density of mean_j exp(X_j):
    it integrates to 1, has mean 100 (E exp(X_j) = 100 by construction of mu, mvn_pdf.f90:19-24) and is non-negative up to
    the truncation ripple."""
    xs = np.linspace(0.0, 300.0, 3001)
    pdf = T.drivers.cos_approximate(xs, _table(), 0.0, 300.0, n_terms=32)
    dx = xs[1] - xs[0]
    trap = lambda f: float(np.sum(0.5 * (f[1:] + f[:-1])) * dx)
    assert abs(trap(pdf) - 1.0) < 1e-6
    assert abs(trap(xs * pdf) - 100.0) < 0.05
    assert pdf.min() > -2e-4 and 0.01 < pdf.max() < 0.03
    p = T.drivers.mvn(4, 64)
    W = T.drivers.chf_weights(p)
    assert W.shape == (32, 4 * 65) and np.array_equal(W, _weights(p.par[:65], p.par[65:130], 4))
    assert np.all(T.drivers.cos_approximate(xs, _table()[:4], 0.0, 300.0, n_terms=8) == 0.0)      # n_terms > size(phis)


def test_oracle_reproduces_reference_chf_table():
    d = 4
    s = O.mvn_setup(d, 64)                       # even N -> 65 like the driver (:48-52)
    r = O.Oracle(s).run(20, piv=1, use_quad=False, use_tru=False)
    nq = int(s.n[0])
    assert nq == 65
    W = _weights(s.par[:nq], s.par[nq:2 * nq], d)
    got = np.array([O.quad_complex(r.cores, w) for w in W])
    tab = _table()
    assert np.abs(got - tab).max() < TOL
    assert abs(got[0] - 1.0) < 1e-6 and got[0].imag == 0.0          # omega = 0: the density integrates to 1


@pytest.mark.gpu
def test_gpu_reproduces_reference_chf_table_and_oracle():
    d = 4
    p = T.drivers.mvn(d, 64)
    t = p.make(use_quad=False, use_tru=False)
    g = t.dmrgg(20, p.accuracy, 1)
    nq = int(p.n[0])
    W = _weights(p.par[:nq], p.par[nq:2 * nq], d)
    got = t.quad_complex(W)
    assert np.abs(got - _table()).max() < TOL
    s = O.mvn_setup(d, 64)
    o = O.Oracle(s).run(20, piv=1, use_quad=False, use_tru=False)
    want = np.array([O.quad_complex(o.cores, w) for w in W])
    assert list(g.ranks) == list(o.ranks)
    # exp-based integrand: device and libm exp differ in the last bit, the quadratures agree far below the table's resolution
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_chf_program_prints_reference_layout(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ttcross_b200", "programs")])
    tab = tmp_path / "chf.txt"
    vals = json.load(open(os.path.join(HERE, "golden", "reference", "chf_table.json")))["values"]
    tab.write_text("".join(f"{a!r} {b!r}\n" for a, b in vals))
    env = dict(os.environ, TTC_CHF_TABLE=str(tab), TTC_SEED="1")
    r = subprocess.run([os.path.join(ROOT, "ttcross_b200", "programs", "bin", "test_crs_chf"), "4", "64", "20", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert re.search(r"quadratur:\s+65\s+quadratur:\s+65 \(adjusted\)", out) and out.rstrip().endswith("Good bye.")
    comp = re.findall(r"computed value:\s+(\S+)\s+(\S+)", out)
    assert len(comp) == 32
    got = np.array([complex(float(a.replace("E", "e")), float(b.replace("E", "e"))) for a, b in comp])
    assert np.abs(got - _table()).max() < TOL
    digits = [float(x) for x in re.findall(r"correct digits:\s*(\S+)", out)]
    assert len(digits) == 32 and min(digits[:6]) > 4.0               # the leading frequencies carry >4 digits of the table


@pytest.mark.gpu
def test_pdf_program_writes_the_cos_density(tmp_path):
    """test_crs_pdf.f90: chf pipeline + cos_approximate_array on 200 points of [0, 300], two es25.17 columns."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ttcross_b200", "programs")])
    out = tmp_path / "pdf.txt"
    env = dict(os.environ, TTC_PDF_OUT=str(out), TTC_SEED="1", TTC_QUIET="1")
    r = subprocess.run([os.path.join(ROOT, "ttcross_b200", "programs", "bin", "test_crs_pdf"), "4", "64", "20", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "Phi values computed." in r.stdout and "Writing PDF output to:" in r.stdout, r.stdout + r.stderr
    rows = np.loadtxt(out)
    assert rows.shape == (200, 2) and rows[0, 0] == 0.0 and rows[-1, 0] == 300.0
    # the same pipeline through the Python mirror
    p = T.drivers.mvn(4, 64)
    t = p.make(use_quad=False, use_tru=False); t.set_seed(1)
    t.dmrgg(20, p.accuracy, 1)
    phis = t.quad_complex(T.drivers.chf_weights(p))
    want = T.drivers.cos_approximate(rows[:, 0], phis, 0.0, 300.0, n_terms=32)
    np.testing.assert_allclose(rows[:, 1], want, rtol=1e-9, atol=1e-13)
    dx = rows[1, 0] - rows[0, 0]
    assert abs(float(np.sum(0.5 * (rows[1:, 1] + rows[:-1, 1])) * dx) - 1.0) < 1e-4      # a density
