"""Host plumbing for the multi-GPU sweep: one process per GPU, `torch.distributed` only to hand the NCCL unique id
of the library's own communicator from rank 0 to the others (the reference's `mpi_init` + `MPI_COMM_WORLD`,
test_crs_ising.f90:31-36).  The data path (pivot tape, boundary fibers, inverse hand-off, quadrature chain;
lib/dmrgg.f90:763-959, 1209-1246, 1355-1405) runs inside the library over NCCL — nothing here touches it.
"""
from __future__ import annotations

import numpy as np

from . import api


def block_of(nparts: int, nranks: int, rank: int):
    """Partitions [first, last) that rank `rank` of `nranks` runs (block map, floor(nparts*g/nranks))."""
    return (nparts * rank) // nranks, (nparts * (rank + 1)) // nranks


def core_block(own, nparts: int, nranks: int, rank: int, d: int):
    """Cores [first, last] (1-based, inclusive) a rank finalises and returns: own(first partition) .. own(next block)-1,
    the last rank also core d (the dtt_lua / dtt_quad ownership, lib/dmrgg.f90:1248-1257, 1323-1345)."""
    v0, v1 = block_of(nparts, nranks, rank)
    return int(own[v0]), (d if v1 == nparts else int(own[v1]) - 1)


def share(first: int, last: int, nproc: int):
    """lib/default.f90:80-97 through the library's helper."""
    import ctypes as C
    L = api.load_library()
    own = np.zeros(nproc + 1, dtype=np.int32)
    L.ttc_share(first, last, nproc, own.ctypes.data_as(C.POINTER(C.c_int)))
    return own


def broadcast_unique_id(dist, make_id=api.TTCross.comm_unique_id, src: int = 0) -> bytes:
    """Rank `src` creates the id, everybody receives it.  Works on any torch.distributed backend (gloo on CPU, nccl)."""
    box = [make_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise api.TTCrossError(-1, "unique id broadcast failed")
    return bytes(uid)


def attach(t: "api.TTCross", dist) -> None:
    """Give handle `t` the communicator of the current torch.distributed world (collective)."""
    t.comm_init(dist.get_world_size(), dist.get_rank(), broadcast_unique_id(dist))


def gather_cores(t: "api.TTCross", dist, dst: int = 0):
    """All cores on rank `dst` (list, core order); None elsewhere.  Convenience for tests: the reference leaves
    every rank with its own cores only."""
    lo, hi = t.core_range()
    mine = (lo, [np.ascontiguousarray(c) for c in t.cores()])
    box = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(mine, box, dst=dst)
    if dist.get_rank() != dst:
        return None
    out = {}
    for first, cs in box:
        for i, c in enumerate(cs):
            out[first + i] = c
    return [out[k] for k in sorted(out)]
