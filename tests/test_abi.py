"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/ttcross_b200.h declares, validates
its arguments like the reference does, agrees bit-for-bit with the oracle on the host-side helpers, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import ttcross_b200 as T
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ttcross_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttc_[a-z0-9_]+)\s*\(", src)) - {"ttc_uniform_cb"})


def test_library_exports_every_declared_symbol():
    L = T.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", T.build.LIB], text=True)
    exported = set(re.findall(r"\bT (ttc_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    for s in syms:
        getattr(L, s)
    assert L.ttc_version() >= 100


def test_library_is_sm100a_and_has_no_torch_dependency():
    out = subprocess.check_output(["ldd", T.build.LIB], text=True)
    assert "torch" not in out and "libcudart" in out
    sass = subprocess.run(["cuobjdump", "-lelf", T.build.LIB], capture_output=True, text=True)
    if sass.returncode == 0:
        assert "sm_100a" in sass.stdout


def test_create_validates_arguments():
    L = T.load_library()
    with pytest.raises(T.TTCrossError):
        T.TTCross(99, [5, 5, 5], np.zeros(11))                    # unknown integrand
    with pytest.raises(T.TTCrossError):
        T.TTCross(T.ISING, [5, 5, 6], np.ones(11))                # Ising needs equal mode sizes
    par = np.ones(11); par[10] = 7
    with pytest.raises(T.TTCrossError) as e:
        T.TTCross(T.ISING, [5, 5, 5], par)                        # 'unknown id' (test_crs_ising.f90:211)
    assert "unknown id" in str(e.value)
    with pytest.raises(T.TTCrossError):
        T.TTCross(T.MVN, [5, 5, 5], np.ones(10), aux=np.ones(3))  # MVN aux too short
    assert L.ttc_last_error(None) != b""


def test_host_helpers_agree_with_oracle_bitwise():
    L = T.load_library()
    for n in (5, 65, 257, 513):
        x, w = T.drivers.lgwt(n)
        xo, wo = O.lgwt(n)
        assert np.array_equal(x, xo) and np.array_equal(w, wo)
    for first, last, nproc in [(1, 8, 8), (1, 8, 3), (1, 63, 8), (1, 4, 2)]:
        own = np.zeros(nproc + 1, dtype=np.int32)
        L.ttc_share(first, last, nproc, own.ctypes.data_as(C.POINTER(C.c_int)))
        assert list(own) == list(O.share(first, last, nproc))
    for seed, v, k in [(1, 0, 0), (1, 0, 1), (5, 3, 12345), (2 ** 63 + 11, 7, 2 ** 40)]:
        assert L.ttc_stream_uniform(seed, v, k) == O.lib().tto_stream_uniform(seed, v, k)


def test_driver_setups_agree_with_oracle_restatement():
    for a, m, n in [("c", 6, 64), ("d", 8, 256), ("e", 6, 512), ("d", 10, 16)]:
        p, s = T.drivers.ising(a, m, n), O.ising_setup(a, m, n)
        assert p.d == s.d and np.array_equal(p.n, s.n) and np.array_equal(p.par, s.par) and np.array_equal(p.quad, s.quad)
        assert p.tru == s.tru and p.accuracy == s.accuracy
    p, s = T.drivers.mvn(8, 128), O.mvn_setup(8, 128)
    assert np.array_equal(p.par, s.par) and np.array_equal(p.aux, s.aux) and np.array_equal(p.quad, s.quad)
    assert p.n[0] == 129                               # N even -> N + 1 (test_crs_mvn.f90:42-45)
    p, s = T.drivers.stdnorm(4, 16), O.stdnorm_setup(4, 16)
    assert np.array_equal(p.par, s.par) and p.accuracy == 5 * 2.220446049250313e-16


def test_results_before_run_are_refused():
    p = T.drivers.ising("c", 4, 8)
    t = p.make()
    r = np.zeros(p.d + 1, dtype=np.int32)
    assert T.load_library().ttc_ranks(t.h, r.ctypes.data_as(C.POINTER(C.c_int))) == 5    # TTC_ERR_STATE
    v = C.c_double()
    assert T.load_library().ttc_quad(t.h, C.byref(v)) == 5


def test_no_gpu_fails_loudly(has_gpu):
    if has_gpu:
        pytest.skip("a CUDA device is present")
    p = T.drivers.ising("c", 4, 8)
    t = p.make()
    with pytest.raises(T.TTCrossError) as e:
        t.dmrgg(4, p.accuracy, 1)
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    # the product path must not import, link or execute anything under oracle/
    pkg = os.path.join(ROOT, "ttcross_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle/", "libttcross_oracle", "tto_"):
                    assert needle not in src, f"{f} references the oracle ({needle})"
    out = subprocess.check_output(["ldd", T.build.LIB], text=True)
    assert "oracle" not in out


def test_closed_form_lottery_matches_literal_loop():
    """The device lottery evaluates lottery2's sequentially accumulated cumulative weights in closed form (piecewise linear
    in exact integer arithmetic).  Here the same code runs on the host and is compared with the literal loop of
    rnd.f90:115-125 (the oracle), including uniforms placed exactly on and next to the cumulative boundaries."""
    L = T.load_library()
    L.ttc_lottery_closed_form.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    rng = np.random.default_rng(1)
    for trial in range(60):
        m = int(rng.integers(2, 40000)) if trial % 3 else int(rng.integers(2, 200))
        nz = int(rng.integers(0, min(64, m - 1) + 1))
        zeros = np.sort(rng.choice(m, nz, replace=False) + 1).astype(np.int32)
        w = np.ones(m)
        w[zeros - 1] = 0
        cnt = 1000
        u = rng.random(cnt)
        scol = m - nz
        Tt = np.zeros(scol + 1)
        for c in range(1, scol + 1):
            Tt[c] = Tt[c - 1] + 1.0 / scol
        pick = rng.integers(1, scol + 1, size=600)
        u[:200] = Tt[pick[:200]]
        u[200:400] = np.nextafter(Tt[pick[200:400]], 0)
        u[400:600] = np.nextafter(Tt[pick[400:600]], 2)
        u[600:604] = [0.0, np.nextafter(1.0, 0), Tt[scol], np.nextafter(Tt[scol], 0)]
        u = np.clip(u, 0, np.nextafter(1.0, 0))
        cells = np.zeros(cnt, dtype=np.int32)
        assert L.ttc_lottery_closed_form(m, zeros.ctypes.data_as(C.POINTER(C.c_int)), nz, u.ctypes.data_as(C.POINTER(C.c_double)),
                                         cnt, cells.ctypes.data_as(C.POINTER(C.c_int))) == 0
        uu = np.concatenate([u, np.zeros(cnt)])
        pts = np.zeros(2 * cnt, dtype=np.int32)
        O.lib().tto_lottery2(cnt, m, 1, O._dp(w), O._dp(np.ones(1)), O._dp(uu), O._ip(pts))
        assert np.array_equal(pts[:cnt], cells), f"trial {trial}: m={m} nz={nz}"
        # the table-free rule of the cluster kernel (one multiplication, literal summation inside the rounding window)
        L.ttc_lottery_fast.argtypes = L.ttc_lottery_closed_form.argtypes
        cells2 = np.zeros(cnt, dtype=np.int32)
        assert L.ttc_lottery_fast(m, zeros.ctypes.data_as(C.POINTER(C.c_int)), nz, u.ctypes.data_as(C.POINTER(C.c_double)),
                                  cnt, cells2.ctypes.data_as(C.POINTER(C.c_int))) == 0
        assert np.array_equal(pts[:cnt], cells2), f"trial {trial} (fast rule): m={m} nz={nz}"


def test_ragged_mode_sizes_are_rejected_for_the_ising_family():
    """type(dtt) carries one mode size per core (tt.f90:18-26) but the drivers always use one quadrature for all modes; the
    Ising integrand here is built on that (weights live at par[n(1) + ind]): unequal sizes are an argument error with a
    message, not a silent misread.  (The nodes-only integrands accept them.)  No GPU needed: ttc_create validates first."""
    import numpy as np
    import pytest
    import ttcross_b200 as T
    p = T.drivers.ising("c", 7, 16)
    p.n = np.array([17, 12, 17, 10, 14, 17], dtype=np.int32)
    p.quad = np.full(int(p.n.sum()), p.quad[0])
    with pytest.raises(T.api.TTCrossError) as e:
        p.make()
    assert "equal mode sizes" in str(e.value)
