"""ctypes wrapper around the CPU oracle (oracle/ttcross_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never from ttcross_b200/.

Also restates the *driver-side* problem setup of the reference programs
(test_crs_ising.f90:39-153, test_crs_mvn.f90:41-133, test_crs_stdnorm.f90:39-131,
lib/mvn_pdf.f90:15-60,85-111) so that tests can build the same `par` blobs the
reference drivers build.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libttcross_oracle.so")

ISING, STDNORM, MVN, COSCOEF = 1, 4, 5, 6
EPS = 2.220446049250313e-16


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ttcross_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.tto_create.restype = C.c_void_p
        L.tto_create.argtypes = [C.c_int, C.c_int, ip, dp, C.c_long, dp, C.c_long]
        L.tto_destroy.argtypes = [C.c_void_p]
        L.tto_run.restype = C.c_int
        L.tto_run.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, ip, dp, C.c_int, C.c_double,
                              C.c_ulonglong, C.c_int]
        L.tto_nsweeps.restype = C.c_int
        L.tto_nsweeps.argtypes = [C.c_void_p]
        L.tto_neval.restype = C.c_longlong
        L.tto_neval.argtypes = [C.c_void_p]
        L.tto_seconds.restype = C.c_double
        L.tto_seconds.argtypes = [C.c_void_p]
        L.tto_quad_final.restype = C.c_double
        L.tto_quad_final.argtypes = [C.c_void_p]
        L.tto_ranks.argtypes = [C.c_void_p, ip]
        L.tto_series.argtypes = [C.c_void_p, C.c_int, dp]
        L.tto_pivlog_count.restype = C.c_long
        L.tto_pivlog_count.argtypes = [C.c_void_p]
        L.tto_pivlog.argtypes = [C.c_void_p, ip, dp]
        L.tto_core.argtypes = [C.c_void_p, C.c_int, dp]
        L.tto_text.restype = C.c_long
        L.tto_text.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        L.tto_integrand.restype = C.c_double
        L.tto_integrand.argtypes = [C.c_void_p, ip]
        L.tto_lgwt.argtypes = [C.c_int, dp, dp]
        L.tto_share.argtypes = [C.c_int, C.c_int, C.c_int, ip]
        L.tto_lottery2.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, ip]
        L.tto_stream_uniform.restype = C.c_double
        L.tto_stream_uniform.argtypes = [C.c_ulonglong, C.c_int, C.c_ulonglong]
        L.tto_idamax.restype = C.c_int
        L.tto_idamax.argtypes = [C.c_long, dp]
        L.tto_d2_lual.argtypes = [C.c_long, C.c_int, dp, dp, C.c_int]
        L.tto_d2_luar.argtypes = [C.c_long, C.c_int, dp, dp, C.c_int]
        L.tto_qr_thin.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.tto_quad_complex.argtypes = [C.c_int, ip, ip, dp, dp, dp, dp]
        L.tto_tt_svd.argtypes = [C.c_int, ip, ip, dp, C.c_double, C.c_int]
        L.tto_tt_svd.restype = C.c_int
        L.tto_tt_ort.argtypes = [C.c_int, ip, ip, dp]
        L.tto_tt_ort.restype = C.c_int
        L.tto_erank.restype = C.c_double
        L.tto_erank.argtypes = [C.c_int, ip, ip]
        L.tto_fmt_e.restype = C.c_int
        L.tto_fmt_e.argtypes = [C.c_double, C.c_int, C.c_int, C.c_char_p]
        L.tto_num_threads.restype = C.c_int
        L.tto_set_num_threads.argtypes = [C.c_int]
        L.tto_set_exp_mode.argtypes = [C.c_void_p, C.c_int]
        L.tto_set_rank_concurrency.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


# --------------------------------------------------------------------------------------
# driver-side setup (restated from the reference programs)
# --------------------------------------------------------------------------------------
def lgwt(n: int):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().tto_lgwt(n, _dp(x), _dp(w))
    return x, w


# test_crs_ising.f90:71-100 (first 17 significant digits are all a double can hold)
_TPI = 6.2831853071795864769
_LOG2 = 0.69314718055994530942
_ZETA3 = 1.2020569031595942854
_C3 = 0.78130241289648629687
ISING_TRU = {
    ("c", 2): 1.0,
    ("c", 3): _C3,
    ("c", 4): 0.70119986017642999982,
    ("c", 5): 0.66575980019993742832,
    ("c", 6): 0.64863420903100707526,
    ("c", 8): 0.63548402675916322614,
    ("c", 16): 0.63050394617323726351,
    ("c", 32): 0.63047350420733980638,
    ("c", 64): 0.63047350337438679649,
    ("c", 128): 0.63047350337438679612,
    ("c", 256): 0.63047350337438679612,
    ("c", 512): 0.63047350337438679612,
    ("c", 1024): 0.63047350337438679612,
    ("d", 2): 1.0 / 3,
    ("d", 3): 8.0 + _TPI ** 2 / 3 - 27.0 * _C3,
    ("d", 4): _TPI ** 2 / 9.0 - 1.0 / 6 - 7.0 * _ZETA3 / 2,
    ("d", 5): 0.0024846057623403154800,
    ("d", 6): 0.00048914170018803477510,
    ("e", 2): 6.0 - 8.0 * _LOG2,
    ("e", 3): 10.0 - _TPI ** 2 / 2 - 8.0 * _LOG2 + 32.0 * _LOG2 ** 2,
    ("e", 4): 22.0 - 82.0 * _ZETA3 - 24.0 * _LOG2 + 176.0 * _LOG2 ** 2 - 256.0 * _LOG2 ** 3 / 3
    + 4.0 * (_TPI ** 2) * _LOG2 - 11.0 * _TPI ** 2 / 6.0,
    ("e", 5): 0.0034936537117295217407,
    ("e", 6): 0.00068783287182640943700,
}


@dataclass
class Setup:
    kind: int
    d: int
    n: np.ndarray          # int32[d]
    par: np.ndarray        # float64
    aux: np.ndarray        # float64 (MVN: mu | inv_cov | denom) or empty
    quad: np.ndarray       # float64[sum(n)] concatenated weight vectors
    accuracy: float
    tru: float             # 0.0 == absent (Ising), as in the reference driver
    label: str = ""
    extra: dict = field(default_factory=dict)


def ising_setup(a: str, index: int, n: int) -> Setup:
    """test_crs_ising.f90:39-44, 61-69, 102-104, 130-153."""
    a = a.lower()
    m = index
    if n % 2 == 0:
        n += 1
    par = np.zeros(2 * n + 1)
    par[2 * n] = {"c": 1.0, "d": 2.0, "e": 3.0}[a]
    x, w = lgwt(n)
    w = 0.5 * w                       # dscal(n, 0.5d0, par(n+1), 1)
    x = (x + 1.0) / 2                 # [-1,1] -> [0,1]
    rescale = a in ("d", "e") and m >= 10
    val = float(n // 2)               # dble(n/2), integer division
    if rescale:
        w = (5.0 * val) * w
    else:
        w = val * w
    par[:n] = x
    par[n:2 * n] = w
    d = m - 1
    quad = np.full(d * n, 1.0 / val)
    return Setup(ISING, d, np.full(d, n, dtype=np.int32), par, np.zeros(0), quad, 500 * EPS,
                 ISING_TRU.get((a, m), 0.0), f"ising {a} {m} n={n}", {"rescale": rescale})


def _interval_setup(n: int, a: float, b: float):
    if n % 2 == 0:
        n += 1
    x, w = lgwt(n)
    x = 0.5 * ((b - a) * x + (a + b))
    w = (0.5 * (b - a)) * w
    return n, x, w


def stdnorm_setup(d: int, n: int) -> Setup:
    """test_crs_stdnorm.f90:39-131 (acc = 5 eps, domain [-10,10])."""
    n, x, w = _interval_setup(n, -10.0, 10.0)
    par = np.concatenate([x, w])
    tru = math.sqrt(3.141592653589793238) ** d
    return Setup(STDNORM, d, np.full(d, n, dtype=np.int32), par, np.zeros(0), np.tile(w, d), 5 * EPS, tru,
                 f"stdnorm d={d} n={n}")


def _powi(x: float, m: int) -> float:
    """libgcc __powidf2: what gfortran emits for real**integer."""
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
    """lib/mvn_pdf.f90:15-60, 85-111.  LAPACK dgetrf/dgetri -> a deterministic Gauss-Jordan."""
    sigma, corr = 0.4, 0.5
    X0 = math.log(100.0)
    mu = np.full(n, X0 + (r - 0.5 * (sigma * sigma)) * T)
    cov = np.empty((n, n))
    for i in range(n):
        for j in range(n):
            cov[i, j] = (sigma * sigma if i == j else sigma * corr * sigma) * T
    inv, det = _inv_det(cov)
    return mu, inv, det


def mvn_setup(d: int, n: int) -> Setup:
    """test_crs_mvn.f90:41-133.  a, b are single-precision literals in the reference."""
    a = float(np.float32(0.525170))
    b = float(np.float32(8.525170))
    n, x, w = _interval_setup(n, a, b)
    par = np.concatenate([x, w])
    mu, inv, det = mvn_init(d)
    denom = math.sqrt(_powi(2.0 * 3.141592653589793, d) * det)
    aux = np.concatenate([mu, np.asfortranarray(inv).ravel(order="F"), [denom]])
    return Setup(MVN, d, np.full(d, n, dtype=np.int32), par, aux, np.tile(w, d), 500 * EPS, 1.0, f"mvn d={d} n={n}")


# --------------------------------------------------------------------------------------
# the oracle run
# --------------------------------------------------------------------------------------
@dataclass
class OracleResult:
    status: int
    nsweeps: int
    neval: int
    ranks: np.ndarray
    vals: np.ndarray
    nevals: np.ndarray
    amaxs: np.ndarray
    pivotmaxs: np.ndarray
    pivlog: np.ndarray       # int32 [count, 8]: it, vrank, bond, ii, jj, kk, qq, upd
    pivots: np.ndarray       # float64 [count]
    cores: list              # cores[k-1] : ndarray (r(k-1), n(k), r(k)) Fortran order
    quad_final: float
    text: str
    seconds: float


class Oracle:
    def __init__(self, setup: Setup):
        self.s = setup
        self._n = np.ascontiguousarray(setup.n, dtype=np.int32)
        self._par = np.ascontiguousarray(setup.par, dtype=np.float64)
        self._aux = np.ascontiguousarray(setup.aux, dtype=np.float64)
        self.h = lib().tto_create(setup.kind, setup.d, _ip(self._n), _dp(self._par), self._par.size,
                                  _dp(self._aux) if self._aux.size else None, self._aux.size)

    def __del__(self):
        try:
            if self.h:
                lib().tto_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_exp_mode(self, mode: int):
        """1: exp through include/ttc_detexp.h, the deterministic routine the product uses in its parity mode."""
        lib().tto_set_exp_mode(self.h, mode)

    def integrand(self, ind) -> float:
        a = np.ascontiguousarray(ind, dtype=np.int32)
        return lib().tto_integrand(self.h, _ip(a))

    def run(self, maxrank: int, piv: int = 3, P: int = 1, own=None, accuracy=None, use_quad=True, use_tru=True,
            seed: int = 1, verbose: bool = False) -> OracleResult:
        L = lib()
        s = self.s
        acc = s.accuracy if accuracy is None else accuracy
        ownp = None
        if own is not None:
            own_a = np.ascontiguousarray(own, dtype=np.int32)
            ownp = _ip(own_a)
        quad = np.ascontiguousarray(s.quad, dtype=np.float64) if use_quad else None
        has_tru = int(use_tru and s.tru != 0.0)
        st = L.tto_run(self.h, maxrank, acc, piv, P, ownp, _dp(quad) if quad is not None else None, has_tru, s.tru,
                       seed, int(verbose))
        text_len = L.tto_text(self.h, None, 0)
        buf = C.create_string_buffer(text_len + 1)
        L.tto_text(self.h, buf, text_len + 1)
        text = buf.value.decode()
        if st != 0:
            return OracleResult(st, 0, 0, np.zeros(0, np.int32), *(np.zeros(0),) * 4, np.zeros((0, 8), np.int32),
                                np.zeros(0), [], 0.0, text, 0.0)
        ns = L.tto_nsweeps(self.h)
        ranks = np.zeros(s.d + 1, dtype=np.int32)
        L.tto_ranks(self.h, _ip(ranks))
        series = []
        for which in range(4):
            a = np.zeros(ns + 1)
            L.tto_series(self.h, which, _dp(a))
            series.append(a)
        cnt = L.tto_pivlog_count(self.h)
        pl = np.zeros((cnt, 8), dtype=np.int32)
        pv = np.zeros(cnt)
        if cnt:
            L.tto_pivlog(self.h, _ip(pl), _dp(pv))
        cores = []
        for k in range(1, s.d + 1):
            shp = (int(ranks[k - 1]), int(s.n[k - 1]), int(ranks[k]))
            a = np.zeros(shp, order="F")
            L.tto_core(self.h, k, _dp(a))
            cores.append(a)
        return OracleResult(0, ns, L.tto_neval(self.h), ranks, series[0], series[1].astype(np.int64), series[2],
                            series[3], pl, pv, cores, L.tto_quad_final(self.h), text, L.tto_seconds(self.h))


def tt_full(cores) -> np.ndarray:
    """Contract a list of (r0,n,r1) cores into the full tensor (small cases only)."""
    t = cores[0]
    t = t.reshape(t.shape[1], t.shape[2])
    for c in cores[1:]:
        t = np.tensordot(t, c, axes=([-1], [0]))
    return t.reshape(t.shape[:-1])


def share(first: int, last: int, nproc: int) -> np.ndarray:
    own = np.zeros(nproc + 1, dtype=np.int32)
    lib().tto_share(first, last, nproc, _ip(own))
    return own


def qr_thin(a):
    """ort0_d (lib/ort.f90:17-81): thin Householder QR with LAPACK's conventions -> (q, r)."""
    a = np.asfortranarray(a, dtype=np.float64)
    m, n = a.shape
    q = np.zeros((m, n), order="F")
    r = np.zeros((n, n), order="F")
    lib().tto_qr_thin(m, n, _dp(a), _dp(q), _dp(r))
    return q, r


def tt_ort(cores):
    """dtt_ort (lib/tt.f90:130-198) on a list of cores (r0 x n x r1 arrays) -> new list of cores."""
    d = len(cores)
    n = np.array([c.shape[1] for c in cores], dtype=np.int32)
    r = np.array([cores[0].shape[0]] + [c.shape[2] for c in cores], dtype=np.int32)
    flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, order="F") for c in cores])
    st = lib().tto_tt_ort(d, _ip(n), _ip(r), _dp(flat))
    if st != 0:
        raise ValueError("tt_ort: an unfolding has fewer rows than columns")
    out, off = [], 0
    for c in cores:
        out.append(flat[off:off + c.size].reshape(c.shape, order="F").copy())
        off += c.size
    return out


def quad_complex(cores, weights):
    """ztt_quad (lib/dmrgg.f90:1418-1523) of real cores against complex rank-1 weights [sum n] -> complex."""
    d = len(cores)
    n = np.array([c.shape[1] for c in cores], dtype=np.int32)
    r = np.array([cores[0].shape[0]] + [c.shape[2] for c in cores], dtype=np.int32)
    flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, order="F") for c in cores])
    w = np.ascontiguousarray(weights, dtype=np.complex128)
    wre, wim = np.ascontiguousarray(w.real), np.ascontiguousarray(w.imag)
    out = np.zeros(2)
    lib().tto_quad_complex(d, _ip(n), _ip(r), _dp(flat), _dp(wre), _dp(wim), _dp(out))
    return complex(out[0], out[1])


def coscoef_setup(d: int, n: int) -> Setup:
    """test_crs_coscoeff.f90:70-186 (no par, no quad in the dtt_dmrgg call)."""
    x0 = math.log(100.0)
    sig = np.full(d, 0.4)
    mean = x0 + (0.0 - 0.5 * sig ** 2) * 1.0
    cov = np.where(np.eye(d, dtype=bool), np.outer(sig, sig) * 1.0, (np.outer(sig, 0.5 * sig)) * 1.0)
    aux = np.concatenate([mean, np.asfortranarray(cov).ravel(order="F"), [0.525170185988090843, 8.52517018598809173]])
    return Setup(COSCOEF, d, np.full(d, n, dtype=np.int32), np.arange(n, dtype=np.float64), aux, np.ones(d * n), 500 * EPS, 0.0,
                 f"coscoeff d={d} n={n}")


def tt_svd(cores, tol=-1.0, rmax=0):
    """dtt_svd (lib/tt.f90:307-368): TT rounding -> new list of cores with the truncated ranks."""
    d = len(cores)
    n = np.array([c.shape[1] for c in cores], dtype=np.int32)
    r = np.array([cores[0].shape[0]] + [c.shape[2] for c in cores], dtype=np.int32)
    flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, order="F") for c in cores])
    st = lib().tto_tt_svd(d, _ip(n), _ip(r), _dp(flat), float(tol), int(rmax))
    if st != 0:
        raise ValueError("tt_svd: unsupported shape")
    out, off = [], 0
    for k in range(d):
        sz = int(r[k]) * int(n[k]) * int(r[k + 1])
        out.append(flat[off:off + sz].reshape((int(r[k]), int(n[k]), int(r[k + 1])), order="F").copy())
        off += sz
    return out
