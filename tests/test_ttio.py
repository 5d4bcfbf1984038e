"""TT files in the reference's stream format (dtt_write / dtt_read, reference lib/ttio.f90:10-17, 29-108, 196-296; SURVEY 8(f)
rank 3).  The byte layout is restated here independently with `struct` from the Fortran text (sequence type `tthead`,
unformatted stream access, default little-endian) and checked against the library's writer and reader."""
import os
import struct

import numpy as np
import pytest

import ttcross_b200 as T


def _ref_bytes(l, cores, comment=b" " * 64, tail=(0,) * 6):
    """What `write(u) head; write(u) l,m; write(u) n(l:m),r(l-1:m); write(u) x` produce (ttio.f90:71-78)."""
    d = len(cores)
    m = l + d - 1
    n = [c.shape[1] for c in cores]
    r = [cores[0].shape[0]] + [c.shape[2] for c in cores]
    head = b"TT      " + struct.pack("<2i", 1, 0) + struct.pack("<4i", 2048, 0, 0, 0) + comment + struct.pack("<8i", l, m, *tail)
    assert len(head) == 128
    body = struct.pack("<2i", l, m) + struct.pack(f"<{d}i", *n) + struct.pack(f"<{d + 1}i", *r)
    x = b"".join(np.asarray(c, dtype="<f8").reshape(-1, order="F").tobytes() for c in cores)
    return head + body + x


def _train(seed, n, r):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((r[k], n[k], r[k + 1])) for k in range(len(n))]


@pytest.mark.parametrize("l,n,r", [(1, [5, 4, 6], [1, 3, 4, 1]), (1, [33] * 5, [1, 8, 16, 16, 8, 1]), (3, [2, 7], [2, 5, 3])])
def test_writer_matches_fortran_stream_layout(tmp_path, l, n, r):
    cores = _train(1, n, r)
    p = str(tmp_path / "a.tt")
    T.tt_write(p, cores, l=l)
    assert open(p, "rb").read() == _ref_bytes(l, cores)


def test_reader_accepts_reference_files_and_round_trips(tmp_path):
    cores = _train(2, [9, 8, 7, 6], [1, 4, 6, 5, 1])
    p = str(tmp_path / "ref.tt")
    # a file as the Fortran writer leaves it: uninitialised comment and trailing header integers (ttio.f90:14-16, 72-74)
    open(p, "wb").write(_ref_bytes(1, cores, comment=bytes(range(64)), tail=(7, -3, 99, 0, 1, 2)))
    l, got = T.tt_read(p)
    assert l == 1 and len(got) == 4
    for a, b in zip(got, cores):
        assert a.shape == b.shape and np.array_equal(a, b)
    q = str(tmp_path / "again.tt")
    T.tt_write(q, got)
    l2, got2 = T.tt_read(q)
    assert all(np.array_equal(a, b) for a, b in zip(got2, cores))


def test_reader_rejects_what_the_reference_rejects(tmp_path):
    p = str(tmp_path / "bad.tt")
    good = _ref_bytes(1, _train(3, [3, 3], [1, 2, 1]))
    open(p, "wb").write(b"XX" + good[2:])                      # ttio.f90:236-240 "not TT header in file"
    with pytest.raises(T.TTCrossError, match="not TT header"):
        T.tt_read(p)
    open(p, "wb").write(good[:8] + struct.pack("<i", 2) + good[12:])   # ttio.f90:241-245 version
    with pytest.raises(T.TTCrossError, match="version"):
        T.tt_read(p)
    open(p, "wb").write(good[:-8])                             # truncated cores (ttio.f90:262 err=114)
    with pytest.raises(T.TTCrossError, match="cores"):
        T.tt_read(p)
    with pytest.raises(T.TTCrossError, match="not exist"):
        T.tt_read(str(tmp_path / "missing.tt"))                # ttio.f90:210-214


@pytest.mark.gpu
def test_handle_writes_its_train(tmp_path):
    p = T.drivers.ising("c", 6, 32)
    t = p.make()
    t.dmrgg(10, p.accuracy, 1)
    f = str(tmp_path / "c6.tt")
    t.write(f)
    l, cores = T.tt_read(f)
    assert l == 1 and open(f, "rb").read() == _ref_bytes(1, t.cores())
    for a, b in zip(cores, t.cores()):
        assert np.array_equal(a, b)
