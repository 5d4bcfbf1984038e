"""Multi-process worker of the multi-GPU parity tests (launched by torch.distributed.run, one process per GPU).

Every rank runs the same problem with the same partition through the C-ABI; the library exchanges pivots, boundary
fibers and quadrature chains over NCCL.  Rank 0 compares tape / ranks / neval / values / gathered cores with the CPU
oracle run at the SAME partition (results are a function of the partition, not of the GPU count).

usage: mp_worker.py KIND INDEX N RANK PIV PARTS [mvn|ising|stdnorm] [EXP_MODE]
(EXP_MODE 1: both sides evaluate exp through include/ttc_detexp.h, so the exp-based integrands are compared bit for bit too)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import ttcross_b200 as T
    from parity_util import to_oracle_setup
    from oracle import oracle as O

    kind, index, n, R, piv, parts = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    family = sys.argv[7] if len(sys.argv) > 7 else "ising"
    exp_mode = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # plumbing only: the unique id travels as a Python object
    if family == "ising":
        p = T.drivers.ising(kind, index, n)
    elif family == "mvn":
        p = T.drivers.mvn(index, n)
    else:
        p = T.drivers.stdnorm(index, n)
    exact = family == "ising" or exp_mode == 1
    t = p.make(device=local)
    t.set_partition(parts)
    t.set_exp_mode(exp_mode)
    T.multi.attach(t, dist)
    ok = True
    msg = ""
    for rep in range(2):                     # second run: buffers and communicator reused
        g = t.dmrgg(R, p.accuracy, piv)
        q = t.quad()                         # collective
        cores = T.multi.gather_cores(t, dist, 0)
        lo, hi = t.core_range()
        exp_lo, exp_hi = T.multi.core_block(T.multi.share(1, p.d - 1, parts), parts, world, rank, p.d)
        assert (lo, hi) == (exp_lo, exp_hi), ((lo, hi), (exp_lo, exp_hi))
        if rank == 0:
            orc = O.Oracle(to_oracle_setup(p))
            orc.set_exp_mode(exp_mode)
            o = orc.run(maxrank=R, piv=piv, P=parts, accuracy=p.accuracy, seed=1)
            try:
                assert np.array_equal(g.pivlog, o.pivlog), "pivot tape indices differ"
                assert np.array_equal(g.ranks, o.ranks), (g.ranks, o.ranks)
                assert g.neval == o.neval and np.array_equal(g.nevals, o.nevals), (g.neval, o.neval)
                assert g.nsweeps == o.nsweeps
                if exact:
                    assert np.array_equal(g.pivots, o.pivots), "pivot values differ"
                    assert np.array_equal(g.vals, o.vals), f"per-sweep values differ {g.vals - o.vals}"
                    assert np.array_equal(g.amaxs, o.amaxs) and np.array_equal(g.pivotmaxs, o.pivotmaxs)
                    assert q == o.quad_final, (q, o.quad_final)
                    for k, (a, b) in enumerate(zip(cores, o.cores), start=1):
                        assert a.shape == b.shape and np.array_equal(a, b), f"core {k} differs"
                else:
                    np.testing.assert_allclose(g.vals, o.vals, rtol=1e-10)
                    np.testing.assert_allclose(q, o.quad_final, rtol=1e-10)
                    for a, b in zip(cores, o.cores):
                        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9 * np.abs(b).max())
            except AssertionError as e:      # keep the other ranks from hanging in the next collective
                ok, msg = False, f"run {rep}: {e}"
        # every rank holds the complete tape and log
        box = [g.pivlog.tobytes() + g.vals.tobytes()]
        dist.broadcast_object_list(box, src=0)
        if box[0] != g.pivlog.tobytes() + g.vals.tobytes():
            ok, msg = False, f"rank {rank}: tape / values differ from rank 0"
        flags = [None] * world
        dist.all_gather_object(flags, (ok, msg))
        if not all(f[0] for f in flags):
            if rank == 0:
                print("MP PARITY FAILED:", [f[1] for f in flags if not f[0]])
            dist.destroy_process_group()
            sys.exit(1)
    if rank == 0:
        print(f"MP PARITY OK world={world} parts={parts} ranks={list(g.ranks)} neval={g.neval} val={g.vals[-1]!r} ms={g.device_ms:.3f}")
    t.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
