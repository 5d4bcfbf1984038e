"""The C++ twins of the reference driver programs (ttcross_b200/programs/, built by __graft_entry__.build()): same positional
CLI, banner, per-sweep lines and closing lines as test_crs_ising.f90 / test_crs_mvn.f90 / test_crs_stdnorm.f90."""
import os
import re
import subprocess

import numpy as np
import pytest

import ttcross_b200 as T
from parity_util import to_oracle_setup
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ttcross_b200", "programs", "bin")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ttcross_b200", "programs")])


def test_programs_build_and_fail_loudly_without_gpu(has_gpu):
    T.load_library()
    _build()
    for prog in ("test_crs_ising", "test_crs_mvn", "test_crs_stdnorm", "test_crs_chf", "test_crs_pdf", "test_crs_store"):
        assert os.access(os.path.join(BIN, prog), os.X_OK)
    if not has_gpu:
        r = subprocess.run([os.path.join(BIN, "test_crs_ising"), "c", "4", "8", "4", "1"], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)     # the reference's write(*,*) msg; stop


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [1, 4])
def test_ising_program_matches_oracle_log(parts):
    _build()
    env = dict(os.environ, TTC_PARTITIONS=str(parts), TTC_SEED="1")
    r = subprocess.run([os.path.join(BIN, "test_crs_ising"), "c", "6", "64", "16", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "Hi, this is TT cross interpolation computing Ising integral..." in out and out.rstrip().endswith("Good bye.")
    assert re.search(r"quadratur:\s+65 \(adjusted\)", out) and re.search(r"TT ranks :\s+16", out) and re.search(r"MPI procs:\s+%d" % parts, out)
    p = T.drivers.ising("c", 6, 64)
    o = O.Oracle(to_oracle_setup(p)).run(maxrank=16, piv=1, P=parts, accuracy=p.accuracy, seed=1)
    strip = lambda x: re.sub(r"time: \S+", "time: *", x)
    if parts > 1:      # rank 0's printed effective rank runs ahead of the reference's hop-per-sweep tape (cosmetic)
        strip = lambda x: re.sub(r"rank\s*\S+", "rank *", re.sub(r"time: \S+", "time: *", x))
    sweep_lines = [strip(l) for l in out.split("\n") if re.match(r"\s*\d+(::|>>|<<) rank", l)]
    assert sweep_lines == [strip(l) for l in o.text.rstrip("\n").split("\n")]          # identical apart from the time field
    m = re.search(r"computed value:\s*(\S+)", out)
    assert abs(float(m.group(1).replace("E", "e")) / o.quad_final - 1) < 1e-14
    m = re.search(r"\.\.\.with\s+(\d+) evaluations", out)
    assert int(m.group(1)) == o.neval
    assert re.search(r"correct digits:\s+\d+\.\d\d", out)


@pytest.mark.gpu
def test_stdnorm_and_mvn_programs_run():
    _build()
    r = subprocess.run([os.path.join(BIN, "test_crs_stdnorm")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Good bye." in r.stdout, r.stdout + r.stderr
    r = subprocess.run([os.path.join(BIN, "test_crs_mvn"), "6", "16", "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Good bye." in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_store_program_writes_the_train(tmp_path):
    """test_crs_store.f90:129-136: the MVN cross of the pdf pipeline, then the train goes to a file (HDF5 in the reference; the TT
    stream format of lib/ttio.f90 here).  The stored cores must be the ones the same cross leaves in the Python mirror, and
    the rest of the pipeline (phis, COS density) must still run."""
    _build()
    tt_out, pdf_out = tmp_path / "tensor_train.tt", tmp_path / "pdf.txt"
    env = dict(os.environ, TTC_TT_OUT=str(tt_out), TTC_PDF_OUT=str(pdf_out), TTC_SEED="1", TTC_QUIET="1")
    r = subprocess.run([os.path.join(BIN, "test_crs_store"), "4", "32", "12", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "Train written to:" in r.stdout and "Writing PDF output to:" in r.stdout, r.stdout + r.stderr
    l, cores = T.tt_read(str(tt_out))
    p = T.drivers.mvn(4, 32)
    t = p.make(use_quad=False, use_tru=False); t.set_seed(1)
    t.dmrgg(12, p.accuracy, 1)
    want = t.cores()
    assert l == 1 and len(cores) == len(want)
    for a, b in zip(cores, want):       # (the C++ driver and the Python mirror build par / the MVN matrix with their own host arithmetic)
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * float(np.abs(b).max()))
    assert np.loadtxt(pdf_out).shape == (200, 2)
