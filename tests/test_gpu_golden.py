"""The CUDA path (through the C-ABI) against the committed golden fixtures of tests/golden/*.json — no oracle run involved:
ranks, evaluation counts, the pivot tape (full or strided sample + checksum of all records), per-sweep values and the
final quadrature, bit for bit for the Ising integrands (exp-based integrands: indices of the first records + values to 1e-9)."""
import glob
import json
import os

import numpy as np
import pytest

import ttcross_b200 as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _checksum(pl):
    a = pl.astype("int64")
    return int((a * (1 + (abs(a).cumsum(axis=0) % 1000003))).sum() % (2 ** 61 - 1))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.json"))), ids=lambda p: os.path.basename(p)[:-5])
def test_cuda_path_matches_golden(path):
    g = json.load(open(path))
    spec = g["spec"]
    p = T.drivers.ising(*spec[1:]) if spec[0] == "ising" else (T.drivers.mvn(*spec[1:]) if spec[0] == "mvn" else T.drivers.stdnorm(*spec[1:]))
    t = p.make()
    t.set_partition(g["P"])
    t.set_seed(g["seed"])
    r = t.dmrgg(g["maxrank"], p.accuracy, g["piv"])
    assert r.nsweeps == g["nsweeps"] and [int(x) for x in r.ranks] == g["ranks"]
    assert r.neval == g["neval"] and [int(x) for x in r.nevals] == g["nevals"]
    st = g["pivlog_stride"]
    assert len(r.pivlog) == g["npiv"]
    if spec[0] == "ising":
        assert r.pivlog[::st].tolist() == g["pivlog"] and _checksum(r.pivlog) == g["pivlog_checksum"]
        assert [float(v).hex() for v in r.vals] == g["vals_hex"]
        assert float(t.quad()).hex() == g["quad_final_hex"]
        assert [float(v).hex() for v in r.pivots[::st]] == g["pivots_hex"]
    else:
        assert r.pivlog[::st][:, :3].tolist() == [x[:3] for x in g["pivlog"]]
        np.testing.assert_allclose(r.vals, [float.fromhex(v) for v in g["vals_hex"]], rtol=1e-9)
