"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same inputs.
Bar: bit-exact pivot tape / ranks / neval / per-sweep values / cores for the Ising (+-*/) integrands."""
import numpy as np
import pytest

import ttcross_b200 as T
from parity_util import run_both, assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,index,n,R,piv", [
    ("c", 4, 8, 6, 1),
    ("c", 6, 16, 8, 1),
    ("c", 6, 64, 16, 1),      # BASELINE config A
    ("d", 5, 16, 8, 2),
    ("e", 6, 32, 10, 3),
    ("c", 5, 16, 8, 0),
    ("c", 5, 12, 6, -1),
    ("d", 4, 10, 6, -1),
])
def test_ising_single_partition(kind, index, n, R, piv):
    p = T.drivers.ising(kind, index, n)
    t, g, o = run_both(p, R, piv)
    assert_parity(t, g, o, exact=True)


@pytest.mark.parametrize("kind,index,n,R,piv,P", [
    ("c", 6, 16, 8, 1, 2),
    ("c", 6, 64, 16, 1, 4),
    ("c", 10, 32, 10, 2, 3),
    ("c", 10, 32, 10, 2, 8),
    ("e", 6, 16, 8, 2, 2),
    ("c", 8, 12, 6, -1, 3),
    ("c", 8, 16, 8, 0, 4),
])
def test_ising_partitions(kind, index, n, R, piv, P):
    p = T.drivers.ising(kind, index, n)
    t, g, o = run_both(p, R, piv, P=P)
    assert_parity(t, g, o, exact=True)


def test_seed_changes_pivots_but_not_convergence():
    p = T.drivers.ising("c", 6, 32)
    _, g1, o1 = run_both(p, 10, 1, seed=1)
    _, g2, o2 = run_both(p, 10, 1, seed=7)
    assert np.array_equal(g1.pivlog, o1.pivlog) and np.array_equal(g2.pivlog, o2.pivlog)
    assert abs(g1.vals[-1] / g2.vals[-1] - 1) < 1e-6


def test_stdnorm_reject_path():
    p = T.drivers.stdnorm(4, 16)
    t, g, o = run_both(p, 10, 1)
    assert_parity(t, g, o, exact=False, rtol=1e-12)
    assert list(g.ranks) == [1, 1, 1, 1, 1]


@pytest.mark.parametrize("d,n,R,piv,P", [(3, 32, 8, -1, 1), (6, 16, 6, 1, 1), (6, 16, 6, 1, 2)])
def test_mvn(d, n, R, piv, P):
    p = T.drivers.mvn(d, n)
    t, g, o = run_both(p, R, piv, P=P)
    assert_parity(t, g, o, exact=False, rtol=1e-10)


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("kind,index,n,R,piv,P", [("c", 6, 64, 16, 1, 1), ("d", 6, 16, 8, 2, 2), ("c", 10, 32, 10, 2, 8)])
def test_lottery_modes_agree(mode, kind, index, n, R, piv, P):
    """mode 1 draws the lottery on the host with the literal loop of rnd.f90; mode 0/2 draw it on the device in closed form."""
    p = T.drivers.ising(kind, index, n)
    t0 = p.make(); t0.set_partition(P); g0 = t0.dmrgg(R, p.accuracy, piv)
    t1 = p.make(); t1.set_partition(P); t1.set_lottery_mode(mode); g1 = t1.dmrgg(R, p.accuracy, piv)
    assert np.array_equal(g0.pivlog, g1.pivlog) and np.array_equal(g0.pivots, g1.pivots)
    assert np.array_equal(g0.vals, g1.vals) and g0.neval == g1.neval
    assert g0.text.split("time:")[0] == g1.text.split("time:")[0]


@pytest.mark.parametrize("P", [1, 3])
def test_accuracy_exit_and_log_format(P):
    # loose accuracy: three consecutive sweeps with pivotmax <= accuracy * amax end the run (dmrgg.f90:1012-1019); the exit
    # falls in the middle of a CUDA graph and, with partitions, the last quadrature still runs on the second stream
    p = T.drivers.ising("c", 6, 32)
    t, g, o = run_both(p, 30, 1, P=P, accuracy=1e-6)
    assert_parity(t, g, o, exact=True)
    assert g.nsweeps < 29
    lines_g, lines_o = g.text.strip().split("\n"), o.text.strip().split("\n")
    assert len(lines_g) == len(lines_o) == g.nsweeps + 1
    import re
    strip = lambda x: re.sub(r"time: \S+", "time: *", x)
    if P > 1:      # rank 0's printed effective rank runs ahead of the reference's hop-per-sweep tape (cosmetic)
        strip = lambda x: re.sub(r"rank\s*\S+", "rank *", re.sub(r"time: \S+", "time: *", x))
    for a, b in zip(lines_g, lines_o):
        assert strip(a) == strip(b), (a, b)       # identical apart from the time field


def test_repeated_runs_reuse_buffers_and_agree():
    p = T.drivers.ising("c", 6, 32)
    t = p.make(); t.set_partition(2)
    g1 = t.dmrgg(10, p.accuracy, 2)
    c1 = t.cores()
    g2 = t.dmrgg(10, p.accuracy, 2)
    assert np.array_equal(g1.pivlog, g2.pivlog) and np.array_equal(g1.vals, g2.vals)
    for a, b in zip(c1, t.cores()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("kind,index,n,R,P", [("c", 6, 20, 8, 1), ("c", 7, 12, 6, 2), ("d", 5, 12, 6, 1), ("e", 5, 10, 5, 1)])
def test_superblock_kernels_agree(kind, index, n, R, P):
    """The tiled superblock kernel (per-row / per-column partial recurrences, register-blocked residual) against the plain
    one-thread-per-element kernel on the same device state: both argmaxes and both values bit for bit; the stored variant
    too; the DFMA variant (not used by the sweep) within rounding."""
    p = T.drivers.ising(kind, index, n)
    t = p.make(); t.set_partition(P)
    t.dmrgg(R, p.accuracy, 1)
    for bond in range(1, p.d):
        try:
            a = t.superblock_probe(bond, variant=0)
        except T.TTCrossError as e:
            assert "another rank" in e.msg
            continue
        b = t.superblock_probe(bond, variant=1)
        s = t.superblock_probe(bond, store=True, variant=0)
        for key in ("argmax_a", "argmax_b", "a", "b", "count"):
            assert a[key] == b[key] == s[key], (bond, key, a, b, s)
        f = t.superblock_probe(bond, variant=2)
        assert f["argmax_a"] == a["argmax_a"] and f["a"] == a["a"]
        assert abs(f["b"] - a["b"]) <= 1e-12 * max(abs(a["a"]), 1e-300)


@pytest.mark.parametrize("kind,index,n,R", [("c", 6, 20, 8), ("c", 7, 12, 4), ("d", 5, 12, 8)])
def test_superblock_dmma_fast_mode_agrees(kind, index, n, R):
    """Fast mode (ttc_superblock_probe variant 3): the residual through mma.sync.m8n8k4.f64 (DMMA).  Not bit-identical by
    construction (FMA, tensor-core summation order); asserted: the evaluated maximum is the same element and value, the
    residual argmax is the same element unless two candidates tie to rounding, and its value agrees to 1e-12 of |a|max."""
    p = T.drivers.ising(kind, index, n)
    t = p.make()
    t.dmrgg(R, p.accuracy, 1)
    for bond in range(1, p.d):
        a = t.superblock_probe(bond, variant=0)
        f = t.superblock_probe(bond, variant=3)
        assert f["count"] == a["count"] and f["argmax_a"] == a["argmax_a"] and f["a"] == a["a"]
        assert abs(f["b"] - a["b"]) <= 1e-12 * max(abs(a["a"]), 1e-300), (bond, a, f)
        if f["argmax_b"] != a["argmax_b"]:
            assert abs(abs(f["b"]) - abs(a["b"])) <= 1e-13 * max(abs(a["a"]), 1e-300)


@pytest.mark.parametrize("P", [1, 2])
def test_superblock_sweep_moderate_shape_bit_exact(P):
    """A full pivoting = -1 run whose superblocks span many row blocks and column tiles of the tiled kernel (8 x 65 x 65 x 8 =
    270 k elements per bond visit at full rank) against the oracle, bit for bit: tape, ranks, neval, values, cores."""
    p = T.drivers.ising("c", 6, 64)
    t, g, o = run_both(p, 8, -1, P=P)
    assert_parity(t, g, o, exact=True)


@pytest.mark.parametrize("P", [1, 4])
def test_superblock_sweep_config_B_mode_size_bit_exact(P):
    """pivoting = -1 at the MODE SIZE of BASELINE config B (n = 257 nodes, d = 9 cores): every bond visit evaluates the whole
    r x 257 x 257 x r superblock (up to 4.2 M elements at rank 8, 65-70 M evaluations in the run) through the tiled kernel, many
    column tiles and row blocks per visit.  Tape, ranks, neval, per-sweep values and cores against the oracle, bit for bit.
    (The full-rank B shape, 32 x 257 x 257 x 32 per visit, is 2 G evaluations per sweep -- minutes of oracle time; it is covered
    by the timing in bench.py's roofline_superblock and the kernel-against-kernel checks above.)"""
    p = T.drivers.ising("c", 10, 256)
    t, g, o = run_both(p, 8, -1, P=P)
    assert_parity(t, g, o, exact=True)
    assert g.neval > 60_000_000


def test_superblock_kernel_mvn_matches_plain():
    p = T.drivers.mvn(4, 16)
    t = p.make()
    t.dmrgg(6, p.accuracy, 1)
    for bond in range(1, p.d):
        a, b = t.superblock_probe(bond, variant=0), t.superblock_probe(bond, variant=1)
        for key in ("argmax_a", "argmax_b", "a", "b", "count"):
            assert a[key] == b[key], (bond, key, a, b)


def test_uniform_callback_feeds_the_host_lottery():
    """ttc_set_uniform_callback: the caller supplies the uniforms (what a Fortran driver would do with random_number).  Fed with
    the library's own stream it must reproduce the device-lottery run bit for bit; fed with another stream it must differ."""
    p = T.drivers.ising("c", 6, 32)
    t0 = p.make(); t0.set_partition(2); t0.set_seed(9)
    g0 = t0.dmrgg(10, p.accuracy, 2)
    L = T.load_library()
    pos = {}

    def stream(vrank, count):
        k0 = pos.get(vrank, 0)
        pos[vrank] = k0 + count
        return [L.ttc_stream_uniform(9, vrank, k0 + i) for i in range(count)]
    t1 = p.make(); t1.set_partition(2); t1.set_uniform_source(stream)
    g1 = t1.dmrgg(10, p.accuracy, 2)
    assert np.array_equal(g0.pivlog, g1.pivlog) and np.array_equal(g0.pivots, g1.pivots) and np.array_equal(g0.vals, g1.vals)
    assert g0.neval == g1.neval and sum(pos.values()) > 0
    rng = np.random.default_rng(0)
    t2 = p.make(); t2.set_partition(2); t2.set_uniform_source(lambda v, c: rng.random(c))
    g2 = t2.dmrgg(10, p.accuracy, 2)
    assert g2.neval != g0.neval or not np.array_equal(g0.pivlog, g2.pivlog)     # another stream: other candidates (the rook search may still meet the same pivots)
    assert abs(g2.vals[-1] / g0.vals[-1] - 1) < 1e-6


@pytest.mark.parametrize("kind,index,n,R,piv,P", [("c", 8, 32, 12, 2, 4), ("e", 6, 24, 10, 1, 1), ("d", 7, 16, 9, 3, 3)])
def test_sweep_schedules_agree(kind, index, n, R, piv, P, monkeypatch):
    """overlapped quadrature + fused exchange/close (default), quadrature in line, separate exchange kernels, the
    wavefront finalisation kernels instead of k_lua_fused: the same pivots, per-sweep values, cores and integral, bit for bit,
    and all equal to the oracle."""
    p = T.drivers.ising(kind, index, n)
    t, g, o = run_both(p, R, piv, P=P)
    assert_parity(t, g, o, exact=True)
    for env in ({"TTC_QUAD_OVERLAP": "0"}, {"TTC_FUSED_SWEEP": "0"}, {"TTC_QUAD_OVERLAP": "0", "TTC_FUSED_SWEEP": "0"},
                {"TTC_GRAPH_SWEEPS": "2"}, {"TTC_NO_LUA_FUSED": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        t2 = p.make(); t2.set_partition(P); t2.set_seed(1)
        g2 = t2.dmrgg(R, p.accuracy, piv)
        for k in env:
            monkeypatch.delenv(k)
        assert np.array_equal(g.pivlog, g2.pivlog) and np.array_equal(g.pivots, g2.pivots), env
        assert np.array_equal(g.vals, g2.vals) and np.array_equal(g.nevals, g2.nevals), env
        assert t.quad() == t2.quad()
        for a, b in zip(t.cores(), t2.cores()):
            assert np.array_equal(a, b)


def test_handles_with_different_rank_capacities_alternate():
    """The opt-in to large dynamic shared memory is a per-kernel cap shared by every handle of the process: a handle with a small
    rank capacity must not lower what a handle with a large one asked for (optin_smem in ttc_engine.cu is monotonic).  Before
    that, the third run below failed with 'invalid argument'."""
    p = T.drivers.ising("c", 6, 24)
    big, small = p.make(), p.make()
    g1 = big.dmrgg(64, p.accuracy, 2)
    gs = small.dmrgg(6, p.accuracy, 2)
    g2 = big.dmrgg(64, p.accuracy, 2)
    assert gs.ranks.max() <= 6
    assert np.array_equal(g1.pivlog, g2.pivlog) and np.array_equal(g1.vals, g2.vals) and np.array_equal(g1.ranks, g2.ranks)
    for a, b in zip(big.cores(), [c.copy() for c in big.cores()]):
        assert np.array_equal(a, b)


def test_no_quadrature_run_with_partitions():
    # dtt_dmrgg without quad= (test_crs_chf.f90:122): no quadrature group at all, the fused kernels still close the sweeps
    p = T.drivers.ising("c", 8, 24)
    t, g, o = run_both(p, 10, 2, P=4, use_quad=False, use_tru=False)
    assert_parity(t, g, o, exact=True)



@pytest.mark.parametrize("family,P,piv", [("stdnorm", 1, 2), ("stdnorm", 2, 1), ("mvn", 1, 1), ("mvn", 3, 1)])
def test_ragged_mode_sizes_nodes_only_integrands(family, P, piv):
    """type(dtt) carries one mode size per core (tt.f90:18-26).  No reference driver uses unequal sizes, but the engine
    (dmrgg.f90) is written for them; the nodes-only integrands read node x(ind) of a common node vector, so core k simply uses
    its first n(k) nodes.  GPU against the oracle at ragged n(k), with the same exp on both sides (parity mode), bit for bit."""
    d = 6
    n = np.array([13, 9, 16, 7, 12, 10], dtype=np.int32)
    base = T.drivers.stdnorm(d, 16) if family == "stdnorm" else T.drivers.mvn(d, 16)
    nmax = int(n.max())
    assert base.n[0] >= nmax
    w = base.quad[:int(base.n[0])]
    quad = np.concatenate([w[:nk] for nk in n])
    p = T.drivers.Problem(base.kind, d, n, base.par, base.aux, quad, base.accuracy, 0.0, f"ragged {family}")
    t, g, o = run_both(p, 6, piv, P=P, exp_mode=1, use_tru=False)
    assert_parity(t, g, o, exact=True)


@pytest.mark.parametrize("kind,index,n,R,piv,P", [("c", 6, 64, 16, 1, 4), ("d", 6, 16, 8, 2, 1), ("c", 10, 32, 10, 2, 8)])
def test_bound_cores_arrive_inside_dmrgg(kind, index, n, R, piv, P):
    """ttc_bind_cores: the reference returns the train inside the call (`arg` inout, lib/dmrgg.f90:11-26).  The bound buffer
    must hold, on return of dmrgg, exactly what ttc_cores delivers -- bit for bit, run after run, also after the handle has
    run an unbound cross in between; a buffer that is too small is refused with the results still available."""
    p = T.drivers.ising(kind, index, n)
    t = p.make()
    t.set_partition(P)
    g0 = t.dmrgg(R, p.accuracy, piv)
    ref = np.concatenate([c.ravel(order="F") for c in t.cores()])
    buf = np.full(t.cores_capacity(R), np.nan)
    t.bind_cores(buf)
    for _ in range(2):
        buf[:] = np.nan
        g = t.dmrgg(R, p.accuracy, piv)
        assert np.array_equal(g.ranks, g0.ranks)
        assert np.array_equal(buf[:ref.size], ref)                  # delivered by dmrgg itself
        assert np.all(np.isnan(buf[ref.size:]))                     # and nothing written past the cores
        views = t.cores(out=buf)                                    # same pointer: views only
        assert np.array_equal(np.concatenate([c.ravel(order="F") for c in views]), ref)
    other = np.empty(ref.size)
    assert np.array_equal(np.concatenate([c.ravel(order="F") for c in t.cores(out=other)]), ref)   # unbound buffer: the usual copy
    small = np.zeros(max(1, ref.size // 2))
    t.bind_cores(small)
    with pytest.raises(T.TTCrossError):
        t.dmrgg(R, p.accuracy, piv)
    assert np.array_equal(np.concatenate([c.ravel(order="F") for c in t.cores()]), ref)
    t.bind_cores(None)
    t.dmrgg(R, p.accuracy, piv)
    assert np.array_equal(np.concatenate([c.ravel(order="F") for c in t.cores()]), ref)
    t.close()
