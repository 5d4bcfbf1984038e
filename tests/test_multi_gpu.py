"""Core blocks over several GPUs (one process per GPU, NCCL): the launcher side.

The CPU tests cover the host logic with a world_size-2 gloo group (block maps, core ownership, the unique-id hand-off,
loud failure without a device).  The GPU test launches tests/mp_worker.py under torch.distributed.run on 2 GPUs, where
rank 0 checks the tape / ranks / neval / values / cores bit for bit against the CPU oracle at the same partition."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ttcross_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_maps_cover_partitions_and_cores():
    for d in (5, 9, 64):
        for parts in range(1, d):
            own = T.multi.share(1, d - 1, parts)
            assert own[0] == 1 and own[parts] == d
            for nranks in range(1, parts + 1):
                seen_v, seen_c = [], []
                for g in range(nranks):
                    v0, v1 = T.multi.block_of(parts, nranks, g)
                    assert v1 > v0                      # every process runs at least one partition
                    seen_v += list(range(v0, v1))
                    lo, hi = T.multi.core_block(own, parts, nranks, g, d)
                    seen_c += list(range(lo, hi + 1))
                assert seen_v == list(range(parts))
                assert seen_c == list(range(1, d + 1))  # every core finalised by exactly one process


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = T.multi.broadcast_unique_id(dist)
        p = T.drivers.ising("c", 6, 16)
        t = p.make()
        t.set_partition(4)
        err = None
        try:
            T.multi.attach(t, dist)                    # collective id hand-off works; the device part must fail loudly here
        except T.TTCrossError as e:
            err = (e.status, e.msg)
        q.put((rank, uid, err))
    finally:
        dist.destroy_process_group()


def test_unique_id_handoff_over_gloo_world2(has_gpu):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] and len(got[0][1]) == 128       # both ranks hold rank 0's id
    if not has_gpu:
        for _, _, err in got:
            assert err is not None and err[0] == 4 and "no CPU fallback" in err[1]   # TTC_ERR_CUDA, loudly


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["c 6 16 8 1 2", "c 10 32 10 2 8", "c 10 32 10 2 3", "e 6 16 8 2 4", "c 8 12 6 -1 3", "x 6 16 6 1 2 mvn",
                                 "x 12 16 6 1 5 mvn 1", "x 8 17 6 2 4 stdnorm 1"])
def test_two_gpus_match_oracle_at_same_partition(cfg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = 29700 + (hash(cfg) % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py")] + cfg.split()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MP PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
