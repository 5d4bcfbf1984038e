"""Host mirror of the reference driver programs' problem setup.

test_crs_ising.f90:39-153, test_crs_mvn.f90:41-133, test_crs_stdnorm.f90:39-131, lib/mvn_pdf.f90:15-60.
Only input generation lives here (nodes, weights, par blob, quadrature weights, analytic values); the
sweep itself is the C-ABI library.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import api

EPS = 2.220446049250313e-16
_dp = C.POINTER(C.c_double)


def lgwt(n: int):
    """Gauss-Legendre nodes/weights on [-1,1] (lib/quad.f90:97-131), computed by the library's host helper."""
    L = api.load_library()
    x = np.zeros(n)
    w = np.zeros(n)
    L.ttc_lgwt(n, x.ctypes.data_as(_dp), w.ctypes.data_as(_dp))
    return x, w


_TPI = 6.2831853071795864769
_LOG2 = 0.69314718055994530942
_ZETA3 = 1.2020569031595942854
_C3 = 0.78130241289648629687
# test_crs_ising.f90:71-100 rounded to double
ISING_TRU = {
    ("c", 2): 1.0, ("c", 3): _C3, ("c", 4): 0.70119986017642999982, ("c", 5): 0.66575980019993742832,
    ("c", 6): 0.64863420903100707526, ("c", 8): 0.63548402675916322614, ("c", 16): 0.63050394617323726351,
    ("c", 32): 0.63047350420733980638, ("c", 64): 0.63047350337438679649, ("c", 128): 0.63047350337438679612,
    ("c", 256): 0.63047350337438679612, ("c", 512): 0.63047350337438679612, ("c", 1024): 0.63047350337438679612,
    ("d", 2): 1.0 / 3, ("d", 3): 8.0 + _TPI ** 2 / 3 - 27.0 * _C3, ("d", 4): _TPI ** 2 / 9.0 - 1.0 / 6 - 7.0 * _ZETA3 / 2,
    ("d", 5): 0.0024846057623403154800, ("d", 6): 0.00048914170018803477510,
    ("e", 2): 6.0 - 8.0 * _LOG2, ("e", 3): 10.0 - _TPI ** 2 / 2 - 8.0 * _LOG2 + 32.0 * _LOG2 ** 2,
    ("e", 4): 22.0 - 82.0 * _ZETA3 - 24.0 * _LOG2 + 176.0 * _LOG2 ** 2 - 256.0 * _LOG2 ** 3 / 3
    + 4.0 * (_TPI ** 2) * _LOG2 - 11.0 * _TPI ** 2 / 6.0,
    ("e", 5): 0.0034936537117295217407, ("e", 6): 0.00068783287182640943700,
}


@dataclass
class Problem:
    kind: int
    d: int
    n: np.ndarray
    par: np.ndarray
    aux: np.ndarray
    quad: np.ndarray
    accuracy: float
    tru: float          # 0.0 = absent
    label: str = ""
    extra: dict = field(default_factory=dict)

    def make(self, device: int = 0, use_quad: bool = True, use_tru: bool = True) -> "api.TTCross":
        t = api.TTCross(self.kind, self.n, self.par, self.aux, device=device)
        if use_quad:
            t.set_quad(self.quad)
        if use_tru and self.tru != 0.0:
            t.set_tru(self.tru)
        return t


def ising(a: str, index: int, n: int) -> Problem:
    a = a.lower()
    if a not in ("c", "d", "e"):
        raise ValueError(f"unknown integral type: {a}")
    m = index
    if n % 2 == 0:
        n += 1
    x, w = lgwt(n)
    w = 0.5 * w
    x = (x + 1.0) / 2
    rescale = a in ("d", "e") and m >= 10
    val = float(n // 2)
    w = ((5.0 * val) if rescale else val) * w
    par = np.zeros(2 * n + 1)
    par[:n] = x
    par[n:2 * n] = w
    par[2 * n] = {"c": 1.0, "d": 2.0, "e": 3.0}[a]
    d = m - 1
    return Problem(api.ISING, d, np.full(d, n, dtype=np.int32), par, np.zeros(0), np.full(d * n, 1.0 / val), 500 * EPS,
                   ISING_TRU.get((a, m), 0.0), f"test_crs_ising {a} {m} {n}", {"rescale": rescale})


def _interval(n: int, a: float, b: float):
    if n % 2 == 0:
        n += 1
    x, w = lgwt(n)
    return n, 0.5 * ((b - a) * x + (a + b)), (0.5 * (b - a)) * w


def stdnorm(d: int, n: int) -> Problem:
    n, x, w = _interval(n, -10.0, 10.0)
    return Problem(api.STDNORM, d, np.full(d, n, dtype=np.int32), np.concatenate([x, w]), np.zeros(0), np.tile(w, d), 5 * EPS,
                   math.sqrt(3.141592653589793238) ** d, f"test_crs_stdnorm {d} {n}")


def _powi(x: float, m: int) -> float:
    n = abs(m)
    y = x if n % 2 else 1.0
    n >>= 1
    while n:
        x = x * x
        if n % 2:
            y *= x
        n >>= 1
    return 1.0 / y if m < 0 else y


def _inv_det(a):
    """Gauss-Jordan with partial pivoting, elementwise IEEE operations only (bit-deterministic on every host; the reference
    uses LAPACK dgetrf/dgetri whose bits depend on the linked BLAS, so any correct inverse is admissible: it is input data)."""
    n = a.shape[0]
    m = np.concatenate([np.array(a, dtype=np.float64), np.eye(n)], axis=1)
    det = 1.0
    for k in range(n):
        piv = k + int(np.argmax(np.abs(m[k:, k])))
        if piv != k:
            m[[k, piv]] = m[[piv, k]]
            det = -det
        det = det * m[k, k]
        m[k] = m[k] / m[k, k]
        for i in range(n):
            if i != k and m[i, k] != 0.0:
                m[i] = m[i] - m[i, k] * m[k]
    return m[:, n:].copy(), det


def mvn_init(n: int, r: float = 0.0, T: float = 1.0):
    """lib/mvn_pdf.f90:15-60,85-111 (dgetrf/dgetri -> a deterministic Gauss-Jordan)."""
    sigma, corr = 0.4, 0.5
    mu = np.full(n, math.log(100.0) + (r - 0.5 * (sigma * sigma)) * T)
    cov = np.full((n, n), (sigma * corr * sigma) * T)
    np.fill_diagonal(cov, (sigma * sigma) * T)
    inv, det = _inv_det(cov)
    return mu, inv, det


def mvn(d: int, n: int) -> Problem:
    a = float(np.float32(0.525170))   # single-precision literals in test_crs_mvn.f90:75-76
    b = float(np.float32(8.525170))
    n, x, w = _interval(n, a, b)
    mu, inv, det = mvn_init(d)
    denom = math.sqrt(_powi(2.0 * 3.141592653589793, d) * det)
    aux = np.concatenate([mu, np.asfortranarray(inv).ravel(order="F"), [denom]])
    return Problem(api.MVN, d, np.full(d, n, dtype=np.int32), np.concatenate([x, w]), aux, np.tile(w, d), 500 * EPS, 1.0,
                   f"test_crs_mvn {d} {n}")


def coscoef(d: int, n: int) -> Problem:
    """test_crs_coscoeff.f90:70-186: COS-method coefficient tensor of a d-variate Gaussian (X_0 = ln 100, sigma = 0.4,
    corr = 0.5, rate = 0, T = 1) on [a, b]; dtt_dmrgg is called WITHOUT par and WITHOUT quad there."""
    x0 = math.log(100.0)
    sig = np.full(d, 0.4)
    mean = x0 + (0.0 - 0.5 * sig ** 2) * 1.0
    cov = np.where(np.eye(d, dtype=bool), np.outer(sig, sig) * 1.0, (np.outer(sig, 0.5 * sig)) * 1.0)
    lower, upper = 0.525170185988090843, 8.52517018598809173
    aux = np.concatenate([mean, np.asfortranarray(cov).ravel(order="F"), [lower, upper]])
    return Problem(api.COSCOEF, d, np.full(d, n, dtype=np.int32), np.arange(n, dtype=np.float64), aux, np.ones(d * n), 500 * EPS, 0.0,
                   f"test_crs_coscoeff {d} {n}")


def chf_weights(p: Problem, nfreq: int = 32, span: float = 300.0) -> np.ndarray:
    """Complex rank-1 weight sets of test_crs_chf.f90:153-166 / test_crs_pdf.f90:153-166: set k weighs every mode with
    w_q * exp(i * omega_k * exp(x_q) / d), omega_k = k*pi/span; shape [nfreq, d*n] as ttc_quad_complex takes it."""
    nq = int(p.n[0])
    x, w = p.par[:nq], p.par[nq:2 * nq]
    return np.array([np.tile(w * np.exp(1j * (k * math.pi / span) * np.exp(x) / p.d), p.d) for k in range(nfreq)])


def cos_approximate(xs, phis, lower_bound: float, upper_bound: float, n_terms=None) -> np.ndarray:
    """lib/cos_approx.f90:90-127 (cos_approximate_array): COS-method density from characteristic-function values,
    sum_k' coeff_k cos(omega_k (x - a)), coeff_k = 2/(b-a) Re(phi_k exp(-i omega_k a)), first term halved.  Host-side
    post-processing of the 32 values ttc_quad_complex returns (test_crs_pdf.f90:171-183)."""
    xs = np.asarray(xs, dtype=np.float64)
    phis = np.asarray(phis, dtype=np.complex128)
    n = len(phis) if n_terms is None else int(n_terms)
    if n > len(phis):                                   # the reference prints an error and returns zeros (:107-111)
        print(" Error: n_terms exceeds the size of phis.")
        return np.zeros_like(xs)
    out = np.zeros_like(xs)
    pi_over_bound = 3.1415926535897932384626433832795 / (upper_bound - lower_bound)
    for k in range(n):
        omega = k * pi_over_bound
        coeff = 2.0 / (upper_bound - lower_bound) * (phis[k] * np.exp(-1j * omega * lower_bound)).real
        if k == 0:
            coeff = coeff / 2.0
        out = out + coeff * np.cos(omega * (xs - lower_bound))
    return out
