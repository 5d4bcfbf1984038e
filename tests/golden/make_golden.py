"""Generate the golden fixtures of tests/golden/*.json from the CPU oracle.

The reference (Fortran + MPI + BLAS) cannot be built or imported in the build container and ships no golden vectors
for pivots / ranks / neval, so these fixtures pin the ORACLE (and, through the -m gpu tests, the CUDA path)
against regressions; the oracle itself is pinned to the reference only through the analytic integral values
(test_crs_ising.f90:71-100 etc.).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

CASES = [
    # name, setup, maxrank, piv, P, seed
    ("ising_c6_n64_r16_piv1_P1", ("ising", "c", 6, 64), 16, 1, 1, 1),     # BASELINE config A
    ("ising_c6_n64_r16_piv1_P4", ("ising", "c", 6, 64), 16, 1, 4, 1),
    ("ising_d5_n16_r8_piv2_P1", ("ising", "d", 5, 16), 8, 2, 1, 1),
    ("ising_e6_n32_r10_piv3_P2", ("ising", "e", 6, 32), 10, 3, 2, 5),
    ("ising_c5_n12_r6_pivm1_P1", ("ising", "c", 5, 12), 6, -1, 1, 1),
    ("ising_c8_n16_r8_piv0_P4", ("ising", "c", 8, 16), 8, 0, 4, 1),
    ("ising_c10_n32_r10_piv2_P8", ("ising", "c", 10, 32), 10, 2, 8, 3),
    ("ising_c10_n256_r32_piv2_P8", ("ising", "c", 10, 256), 32, 2, 8, 1),  # BASELINE config B (bench partition)
    ("ising_c10_n256_r32_piv2_P1", ("ising", "c", 10, 256), 32, 2, 1, 1),  # BASELINE config B, one partition
    ("mvn_d6_n16_r6_piv1_P2", ("mvn", 6, 16), 6, 1, 2, 1),
    ("stdnorm_d4_n16_piv1_P1", ("stdnorm", 4, 16), 10, 1, 1, 1),
]


def setup_of(spec):
    if spec[0] == "ising":
        return O.ising_setup(spec[1], spec[2], spec[3])
    if spec[0] == "mvn":
        return O.mvn_setup(spec[1], spec[2])
    return O.stdnorm_setup(spec[1], spec[2])


def main():
    for name, spec, R, piv, P, seed in CASES:
        s = setup_of(spec)
        r = O.Oracle(s).run(maxrank=R, piv=piv, P=P, seed=seed)
        big = len(r.pivlog) > 400
        rec = {
            "name": name, "spec": list(spec), "maxrank": R, "piv": piv, "P": P, "seed": seed,
            "nsweeps": r.nsweeps, "neval": int(r.neval), "ranks": [int(x) for x in r.ranks],
            "nevals": [int(x) for x in r.nevals],
            "vals_hex": [float(v).hex() for v in r.vals],
            "quad_final_hex": float(r.quad_final).hex(),
            "npiv": int(len(r.pivlog)),
            # full tape for small cases; for the big ones a strided sample plus a checksum of all records
            "pivlog": r.pivlog.tolist() if not big else r.pivlog[::16].tolist(),
            "pivlog_stride": 1 if not big else 16,
            "pivlog_checksum": int((r.pivlog.astype("int64") * (1 + (abs(r.pivlog.astype("int64")).cumsum(axis=0) % 1000003))).sum() % (2 ** 61 - 1)),
            "pivots_hex": [float(v).hex() for v in (r.pivots if not big else r.pivots[::16])],
        }
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(name, rec["neval"], rec["ranks"], r.vals[-1])


if __name__ == "__main__":
    main()
