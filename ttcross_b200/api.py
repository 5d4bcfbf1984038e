"""ctypes binding of include/ttcross_b200.h.

`TTCross` mirrors the reference call `dtt_dmrgg(tt, fun, par, maxrank=, accuracy=, pivoting=, neval=, quad=, tru=)`
(lib/dmrgg.f90:11-26) followed by `dtt_quad(tt, qq)` (lib/dmrgg.f90:1261).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import build as _build

ISING, STDNORM, MVN, COSCOEF = 1, 4, 5, 6

_lib = None
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
# void cb(void* ctx, int vrank, int count, double* out): `count` uniforms in [0,1) for virtual rank `vrank`
UNIFORM_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, _dp)


class TTCrossError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"ttcross_b200 status {status}: {msg}")
        self.status = status
        self.msg = msg


def load_library(build_if_missing: bool = True):
    """Load libttcross_b200.so (in-tree).  Fails loudly if it is missing and cannot be built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TTC_LIB_PATH", _build.LIB)      # (kernel-tuning experiments load alternative builds)
    if build_if_missing and path == _build.LIB and _build.needs_build() and os.path.exists(_build.NVCC):
        _build.build()
    if not os.path.exists(path):
        raise TTCrossError(-1, f"{path} is missing: run `python -m ttcross_b200.build` (nvcc, sm_100a); there is no CPU fallback")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.ttc_create.restype = C.c_int
    L.ttc_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, _ip, _dp, C.c_long, _dp, C.c_long]
    L.ttc_destroy.argtypes = [vp]
    L.ttc_last_error.restype = C.c_char_p
    L.ttc_last_error.argtypes = [vp]
    L.ttc_set_device.argtypes = [vp, C.c_int]
    L.ttc_set_partition.argtypes = [vp, C.c_int, _ip]
    L.ttc_set_par.argtypes = [vp, _dp, C.c_long]
    L.ttc_set_quad.argtypes = [vp, _dp]
    L.ttc_set_tru.argtypes = [vp, C.c_int, C.c_double]
    L.ttc_set_seed.argtypes = [vp, C.c_ulonglong]
    L.ttc_set_verbose.argtypes = [vp, C.c_int]
    L.ttc_set_uniform_callback.argtypes = [vp, UNIFORM_CB, C.c_void_p]
    L.ttc_set_lottery_mode.argtypes = [vp, C.c_int]
    L.ttc_set_exp_mode.argtypes = [vp, C.c_int]
    L.ttc_converged.argtypes = [vp]
    L.ttc_set_profile.argtypes = [vp, C.c_int]
    L.ttc_dmrgg.argtypes = [vp, C.c_int, C.c_double, C.c_int]
    L.ttc_ranks.argtypes = [vp, _ip]
    L.ttc_core.argtypes = [vp, C.c_int, _dp]
    L.ttc_cores.argtypes = [vp, _dp, C.c_longlong]
    L.ttc_bind_cores.argtypes = [vp, C.c_void_p, C.c_longlong]
    L.ttc_neval.restype = C.c_longlong
    L.ttc_neval.argtypes = [vp]
    L.ttc_nsweeps.argtypes = [vp]
    L.ttc_seconds.restype = C.c_double
    L.ttc_seconds.argtypes = [vp]
    L.ttc_sweep_series.argtypes = [vp, C.c_int, _dp]
    L.ttc_pivlog_count.restype = C.c_long
    L.ttc_pivlog_count.argtypes = [vp]
    L.ttc_pivlog.argtypes = [vp, _ip, _dp]
    L.ttc_text.restype = C.c_long
    L.ttc_text.argtypes = [vp, C.c_char_p, C.c_long]
    L.ttc_quad.argtypes = [vp, _dp]
    L.ttc_lgwt.argtypes = [C.c_int, _dp, _dp]
    L.ttc_share.argtypes = [C.c_int, C.c_int, C.c_int, _ip]
    L.ttc_stream_uniform.restype = C.c_double
    L.ttc_stream_uniform.argtypes = [C.c_ulonglong, C.c_int, C.c_ulonglong]
    L.ttc_superblock_probe.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), _dp, _dp, C.POINTER(C.c_longlong)]
    L.ttc_superblock_probe_ex.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), _dp, _dp, C.POINTER(C.c_longlong)]
    L.ttc_fiber_probe.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]
    L.ttc_launch_count.restype = C.c_longlong
    L.ttc_launch_count.argtypes = [vp]
    L.ttc_l2_flush.argtypes = [vp, C.c_longlong]
    L.ttc_device_ms.restype = C.c_double
    L.ttc_device_ms.argtypes = [vp]
    L.ttc_sweep_kernel_ms.restype = C.c_double
    L.ttc_sweep_kernel_ms.argtypes = [vp]
    L.ttc_sweep_geometry.argtypes = [vp, _ip, _ip]
    L.ttc_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_longlong), _dp]
    L.ttc_fp64_peak.argtypes = [C.c_int, C.c_int, _dp]
    L.ttc_ort.argtypes = [vp]
    L.ttc_svd.argtypes = [vp, C.c_double, C.c_int]
    L.ttc_values.argtypes = [vp, C.c_longlong, _ip, _dp]
    L.ttc_accchk.argtypes = [vp, C.c_longlong, C.c_ulonglong, _dp, _ip]
    L.ttc_quad_complex.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp]
    L.ttc_write.argtypes = [vp, C.c_char_p]
    L.ttc_tt_write.argtypes = [C.c_char_p, C.c_int, C.c_int, _ip, _ip, _dp]
    L.ttc_tt_read_header.argtypes = [C.c_char_p, _ip, _ip, _ip, _ip, C.c_int, C.POINTER(C.c_longlong)]
    L.ttc_tt_read_cores.argtypes = [C.c_char_p, _dp, C.c_longlong]
    L.ttc_qr_thin.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp]
    L.ttc_set_timeline.argtypes = [vp, C.c_int]
    L.ttc_timeline.restype = C.c_long
    L.ttc_timeline.argtypes = [vp, C.c_long, _ip, C.POINTER(C.c_ulonglong), C.POINTER(C.c_char_p), C.c_int]
    L.ttc_comm_unique_id.argtypes = [C.c_void_p]
    L.ttc_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_void_p]
    L.ttc_comm_rank.argtypes = [vp, _ip, _ip]
    L.ttc_core_range.argtypes = [vp, _ip, _ip]
    L.ttc_version.restype = C.c_int
    _lib = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


@dataclass
class CrossResult:
    nsweeps: int
    neval: int
    ranks: np.ndarray
    vals: np.ndarray
    nevals: np.ndarray
    amaxs: np.ndarray
    pivotmaxs: np.ndarray
    pivlog: np.ndarray       # int32 [count, 8]: it, vrank, bond, ii, jj, kk, qq, upd
    pivots: np.ndarray
    text: str
    seconds: float
    device_ms: float
    launches: int


def fp64_peak(device: int = 0, fma: bool = True) -> float:
    """Measured FP64 ceiling of the device in TFLOP/s (DFMA chains, or separate DMUL + DADD when fma is False)."""
    L = load_library()
    v = C.c_double()
    st = L.ttc_fp64_peak(device, int(fma), C.byref(v))
    if st != 0:
        raise TTCrossError(st, "ttc_fp64_peak failed (no CUDA device?)")
    return v.value


def qr_thin(a, device: int = 0, reps: int = 1):
    """ort0_d (lib/ort.f90:17-81): thin QR of an m x n block on the GPU -> (q, r, ms)."""
    L = load_library()
    a = np.asfortranarray(a, dtype=np.float64)
    m, n = a.shape
    q = np.zeros((m, n), order="F")
    r = np.zeros((n, n), order="F")
    ms = C.c_double()
    st = L.ttc_qr_thin(device, m, n, _d(a), _d(q), _d(r), reps, C.byref(ms))
    if st != 0:
        raise TTCrossError(st, L.ttc_last_error(None).decode())
    return q, r, ms.value


def tt_write(path: str, cores, l: int = 1):
    """dtt_write (lib/ttio.f90:29-108): a list of cores (r0 x n x r1 arrays) to a TT file in the reference's stream format."""
    L = load_library()
    d = len(cores)
    n = np.array([c.shape[1] for c in cores], dtype=np.int32)
    r = np.array([cores[0].shape[0]] + [c.shape[2] for c in cores], dtype=np.int32)
    flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, order="F") for c in cores])
    st = L.ttc_tt_write(path.encode(), l, l + d - 1, _i(n), _i(r), _d(flat))
    if st != 0:
        raise TTCrossError(st, L.ttc_last_error(None).decode())


def tt_read(path: str):
    """dtt_read (lib/ttio.f90:196-296) -> (l, list of cores)."""
    L = load_library()
    l, m, tot = C.c_int(), C.c_int(), C.c_longlong()
    st = L.ttc_tt_read_header(path.encode(), C.byref(l), C.byref(m), None, None, 0, C.byref(tot))
    if st != 0:
        raise TTCrossError(st, L.ttc_last_error(None).decode())
    d = m.value - l.value + 1
    n = np.zeros(d, dtype=np.int32)
    r = np.zeros(d + 1, dtype=np.int32)
    st = L.ttc_tt_read_header(path.encode(), C.byref(l), C.byref(m), _i(n), _i(r), d, C.byref(tot))
    flat = np.zeros(tot.value)
    if st == 0:
        st = L.ttc_tt_read_cores(path.encode(), _d(flat), flat.size)
    if st != 0:
        raise TTCrossError(st, L.ttc_last_error(None).decode())
    cores, off = [], 0
    for k in range(d):
        sz = int(r[k]) * int(n[k]) * int(r[k + 1])
        cores.append(flat[off:off + sz].reshape((int(r[k]), int(n[k]), int(r[k + 1])), order="F"))
        off += sz
    return l.value, cores


class TTCross:
    """One TT-cross problem on one GPU (handle of the C-ABI)."""

    def __init__(self, kind: int, n, par, aux=None, device: int = 0):
        L = load_library()
        self._L = L
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.d = int(self.n.size)
        self._par = np.ascontiguousarray(par, dtype=np.float64)
        self._aux = np.ascontiguousarray(aux if aux is not None else np.zeros(0), dtype=np.float64)
        h = C.c_void_p()
        st = L.ttc_create(C.byref(h), kind, self.d, _i(self.n), _d(self._par), self._par.size,
                          _d(self._aux) if self._aux.size else None, self._aux.size)
        if st != 0:
            raise TTCrossError(st, L.ttc_last_error(None).decode())
        self.h = h
        self._check(L.ttc_set_device(h, device))
        self.kind = kind

    def _check(self, st):
        if st != 0:
            raise TTCrossError(st, self._L.ttc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self._L.ttc_destroy(self.h)
            self.h = None
            self._bound = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- optional arguments of dtt_dmrgg
    def set_partition(self, nparts: int, own=None):
        if own is None:
            self._check(self._L.ttc_set_partition(self.h, nparts, None))
        else:
            a = np.ascontiguousarray(own, dtype=np.int32)
            self._check(self._L.ttc_set_partition(self.h, nparts, _i(a)))

    def set_par(self, par):
        a = np.ascontiguousarray(par, dtype=np.float64)
        self._check(self._L.ttc_set_par(self.h, _d(a), a.size))
        self._par = a

    def set_quad(self, quad):
        if quad is None:
            self._check(self._L.ttc_set_quad(self.h, None))
        else:
            a = np.ascontiguousarray(quad, dtype=np.float64)
            assert a.size == int(self.n.sum())
            self._check(self._L.ttc_set_quad(self.h, _d(a)))

    def set_tru(self, tru):
        self._check(self._L.ttc_set_tru(self.h, 0 if tru is None else 1, 0.0 if tru is None else float(tru)))

    def set_seed(self, seed: int):
        self._check(self._L.ttc_set_seed(self.h, seed))

    def set_uniform_source(self, fn):
        """fn(vrank, count) -> array of `count` uniforms in [0,1): replaces the built-in stream (the reference's unseeded
        random_number, rnd.f90:120); implies the host lottery, one call per bond visit and virtual rank.  None removes it."""
        if fn is None:
            self._ucb = None
            self._check(self._L.ttc_set_uniform_callback(self.h, UNIFORM_CB(0), None))
            return

        def _cb(ctx, vrank, count, out):
            vals = np.asarray(fn(int(vrank), int(count)), dtype=np.float64)
            for i in range(count):
                out[i] = vals[i]
        self._ucb = UNIFORM_CB(_cb)            # keep the trampoline alive
        self._check(self._L.ttc_set_uniform_callback(self.h, self._ucb, None))

    def set_lottery_mode(self, mode: int):
        self._check(self._L.ttc_set_lottery_mode(self.h, mode))

    def sweep_kernel(self):
        """(used, ms, cluster size, threads per CTA) of the persistent sweep kernel in the last dmrgg call."""
        cs, th = C.c_int(0), C.c_int(0)
        used = self._L.ttc_sweep_geometry(self.h, C.byref(cs), C.byref(th))
        return bool(used), float(self._L.ttc_sweep_kernel_ms(self.h)), cs.value, th.value

    def set_exp_mode(self, mode: int):
        """0: platform exp (default); 1: deterministic exp shared with the test oracle (parity mode)."""
        self._check(self._L.ttc_set_exp_mode(self.h, mode))

    @property
    def converged(self) -> bool:
        return bool(self._L.ttc_converged(self.h))

    def set_verbose(self, v: bool):
        self._check(self._L.ttc_set_verbose(self.h, int(v)))

    def set_profile(self, on: bool):
        self._check(self._L.ttc_set_profile(self.h, int(on)))

    # ---- dtt_dmrgg
    def dmrgg(self, maxrank: int = -1, accuracy: float = -1.0, pivoting: int = 3) -> CrossResult:
        L = self._L
        self._check(L.ttc_dmrgg(self.h, maxrank, accuracy, pivoting))
        ns = L.ttc_nsweeps(self.h)
        ranks = np.zeros(self.d + 1, dtype=np.int32)
        self._check(L.ttc_ranks(self.h, _i(ranks)))
        series = []
        for which in range(4):
            a = np.zeros(ns + 1)
            self._check(L.ttc_sweep_series(self.h, which, _d(a)))
            series.append(a)
        cnt = L.ttc_pivlog_count(self.h)
        pl = np.zeros((cnt, 8), dtype=np.int32)
        pv = np.zeros(cnt)
        if cnt:
            self._check(L.ttc_pivlog(self.h, _i(pl), _d(pv)))
        self.ranks = ranks
        return CrossResult(ns, L.ttc_neval(self.h), ranks, series[0], series[1].astype(np.int64), series[2], series[3],
                           pl, pv, self.text(), L.ttc_seconds(self.h), L.ttc_device_ms(self.h), L.ttc_launch_count(self.h))

    def text(self) -> str:
        ln = self._L.ttc_text(self.h, None, 0)
        buf = C.create_string_buffer(ln + 1)
        self._L.ttc_text(self.h, buf, ln + 1)
        return buf.value.decode()

    # ---- several processes, one per GPU (replaces MPI_COMM_WORLD of the reference)
    @staticmethod
    def comm_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        st = L.ttc_comm_unique_id(buf)
        if st != 0:
            raise TTCrossError(st, L.ttc_last_error(None).decode())
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        assert len(uid) == 128
        self._check(self._L.ttc_comm_init(self.h, nranks, rank, C.c_char_p(uid)))

    def core_range(self):
        lo, hi = C.c_int(), C.c_int()
        self._check(self._L.ttc_core_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def core(self, k: int) -> np.ndarray:
        shp = (int(self.ranks[k - 1]), int(self.n[k - 1]), int(self.ranks[k]))
        a = np.zeros(shp, order="F")
        self._check(self._L.ttc_core(self.h, k, _d(a)))
        return a

    def cores(self, out=None):
        """The cores this rank holds (all of them on a single GPU), in core order.  `out`: optional caller-owned flat float64
        buffer (>= the total size) the cores are written into, the returned arrays are views of it."""
        lo, hi = self.core_range()
        sizes = [int(self.ranks[k - 1]) * int(self.n[k - 1]) * int(self.ranks[k]) for k in range(lo, hi + 1)]
        if out is None:
            buf = np.empty(sum(sizes))
        else:
            assert out.dtype == np.float64 and out.ndim == 1 and out.flags.c_contiguous and out.size >= sum(sizes)
            buf = out[:sum(sizes)]
        self._check(self._L.ttc_cores(self.h, _d(buf), buf.size))
        out, off = [], 0
        for k, sz in zip(range(lo, hi + 1), sizes):
            out.append(buf[off:off + sz].reshape((int(self.ranks[k - 1]), int(self.n[k - 1]), int(self.ranks[k])), order="F"))
            off += sz
        return out

    def bind_cores(self, out):
        """Bind a caller-owned flat float64 buffer that every later dmrgg() fills with this rank's cores before it returns --
        the reference's `arg` is an inout argument of dtt_dmrgg (lib/dmrgg.f90:11-26).  `cores(out=<the same array>)` then
        only builds the views.  `out=None` removes the binding.  The array is kept alive by the handle."""
        if out is None:
            self._check(self._L.ttc_bind_cores(self.h, None, 0))
            self._bound = None
            return
        assert out.dtype == np.float64 and out.ndim == 1 and out.flags.c_contiguous
        self._check(self._L.ttc_bind_cores(self.h, out.ctypes.data, out.size))
        self._bound = out

    def cores_capacity(self, maxrank: int) -> int:
        """Doubles that always hold the cores of a run with the given maxrank (for bind_cores)."""
        return sum((1 if k == 1 else maxrank) * int(self.n[k - 1]) * (1 if k == self.d else maxrank) for k in range(1, self.d + 1))

    def quad_complex(self, weights) -> np.ndarray:
        """ztt_quad (lib/dmrgg.f90:1418-1523) for several complex rank-1 weight tensors at once.
        weights: complex array [nsets, n(1)+...+n(d)] -> complex array [nsets]."""
        w = np.ascontiguousarray(weights, dtype=np.complex128)
        if w.ndim == 1:
            w = w[None, :]
        assert w.shape[1] == int(self.n.sum())
        wre, wim = np.ascontiguousarray(w.real), np.ascontiguousarray(w.imag)
        ore, oim = np.zeros(w.shape[0]), np.zeros(w.shape[0])
        self._check(self._L.ttc_quad_complex(self.h, w.shape[0], _d(wre), _d(wim), _d(ore), _d(oim)))
        return ore + 1j * oim

    def values(self, ind) -> np.ndarray:
        """dtt_ijk (lib/tt.f90:630-652) at many multi-indices: ind int array [count, d], 1-based."""
        a = np.ascontiguousarray(ind, dtype=np.int32)
        if a.ndim == 1:
            a = a[None, :]
        assert a.shape[1] == self.d
        out = np.zeros(a.shape[0])
        self._check(self._L.ttc_values(self.h, a.shape[0], _i(a), _d(out)))
        return out

    def accchk(self, nlot: int, seed: int = 1):
        """dtt_accchk (lib/dmrgg.f90:1081-1166) -> dict(einf, efro, ainf, afro, pivot)."""
        out = np.zeros(4)
        piv = np.zeros(self.d, dtype=np.int32)
        self._check(self._L.ttc_accchk(self.h, nlot, seed, _d(out), _i(piv)))
        return {"einf": out[0], "efro": out[1], "ainf": out[2], "afro": out[3], "pivot": piv}

    def write(self, path: str):
        """dtt_write (lib/ttio.f90:29-108) of the train this handle holds."""
        self._check(self._L.ttc_write(self.h, path.encode()))

    # ---- dtt_ort (lib/tt.f90:130-198)
    def ort(self):
        self._check(self._L.ttc_ort(self.h))

    def svd(self, tol: float = -1.0, rmax: int = 0):
        """dtt_svd (lib/tt.f90:307-368): TT rounding; afterwards self.ranks / core() / quad() see the rounded train."""
        self._check(self._L.ttc_svd(self.h, tol, rmax))
        r = np.zeros(self.d + 1, dtype=np.int32)
        self._check(self._L.ttc_ranks(self.h, _i(r)))
        self.ranks = r

    # ---- dtt_quad
    def quad(self) -> float:
        v = C.c_double()
        self._check(self._L.ttc_quad(self.h, C.byref(v)))
        return v.value

    # ---- probes
    def superblock_probe(self, bond: int, store: bool = False, reps: int = 1, variant: int = 0):
        """variant 0: tiled kernel, reference arithmetic; 1: plain kernel; 2: tiled kernel with DFMA residual (not bit-exact);
        3: fast mode, residual through the FP64 tensor-core path (mma.sync.m8n8k4.f64 / DMMA; not bit-exact)."""
        idx = (C.c_longlong * 2)()
        val = (C.c_double * 2)()
        ms = C.c_double()
        cnt = C.c_longlong()
        self._check(self._L.ttc_superblock_probe_ex(self.h, bond, int(store), reps, variant, idx, val, C.byref(ms), C.byref(cnt)))
        return {"argmax_a": idx[0], "argmax_b": idx[1], "a": val[0], "b": val[1], "ms": ms.value, "count": cnt.value}

    def fiber_probe(self, bond: int, isrow: bool, ii: int, jj: int, kk: int, qq: int, reps: int = 1):
        cnt = (int(self.n[bond]) * int(self.ranks[bond + 1])) if isrow else (int(self.ranks[bond - 1]) * int(self.n[bond - 1]))
        f = np.zeros(cnt)
        r = np.zeros(cnt)
        ms = C.c_double()
        self._check(self._L.ttc_fiber_probe(self.h, bond, int(isrow), ii, jj, kk, qq, _d(f), _d(r), reps, C.byref(ms)))
        return f, r, ms.value

    def profile(self):
        cap = 32
        names = (C.c_char_p * cap)()
        launches = (C.c_longlong * cap)()
        ms = (C.c_double * cap)()
        cnt = self._L.ttc_profile(self.h, cap, names, launches, ms)
        return {names[i].decode(): (int(launches[i]), float(ms[i])) for i in range(cnt)}

    def set_timeline(self, on: bool):
        self._check(self._L.ttc_set_timeline(self.h, int(on)))

    def timeline(self):
        """[(kernel name, t_ns)] stamps of the last run, in time order (diagnostic)."""
        cap = 65536
        ids = np.zeros(cap, dtype=np.int32)
        ts = np.zeros(2 * cap, dtype=np.uint64)
        names = (C.c_char_p * 64)()
        n = self._L.ttc_timeline(self.h, 2 * cap, _i(ids), ts.ctypes.data_as(C.POINTER(C.c_ulonglong)), names, 64)
        self.timeline_clocks = {int(t): int(c) for t, c in zip(ts[:n], ts[n:2 * n])}
        nm = [x.decode() if x else "?" for x in names]
        special = {40: "k_visits", 100: "fold_done", 41: "v:staged", 42: "v:lot_setup", 43: "v:lot_eval", 44: "v:lot_fold",
                   45: "v:fiber_eval", 46: "v:fiber_fold", 47: "v:rook_done", 48: "v:nbr_done", 49: "v:append_done", 50: "f:xs_staged", 51: "f:pref_issued", 52: "f:eval_done",
                   53: "f:resid_done", 54: "f:stored", 55: "l:drawn", 56: "l:evaluated", 34: "k_quad_inc", 60: "q:lu_staged", 61: "q:chunk_staged", 62: "q:chunk_summed",
                   63: "q:luar_done", 64: "q:end", 35: "k_superblock_t", 36: "k_sweeps", 37: "k_lua_fused", 70: "s:announced", 71: "s:nbr_ready",
                   72: "s:exchanged", 73: "s:all_ready", 74: "s:closed", 75: "x:staged", 76: "x:corner", 77: "x:chains"}
        out = [(special.get(int(i), nm[i] if i < 64 else "?"), int(t)) for i, t in zip(ids[:n], ts[:n])]
        return sorted(out, key=lambda x: x[1])

    def l2_flush(self, nbytes: int = 256 << 20):
        self._check(self._L.ttc_l2_flush(self.h, nbytes))

    def launch_count(self) -> int:
        return int(self._L.ttc_launch_count(self.h))
