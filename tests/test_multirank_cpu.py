"""CPU tests (gloo, world_size 2) of the host-side multi-rank plumbing that bench.py and tools/mg_check.py use on
the GPU box with NCCL: unique-id broadcast, max/sum over ranks, and that the per-rank z-slabs of the bench problem
are exactly the blocks of the reference's decomposition (blockGrid.hpp:151-170) of the global manufactured problem."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, npglobal, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import pyoracle as po
    D = bench.Dist(rank, world, "cpu")
    # 1. the NCCL unique id of rank 0 reaches every rank unchanged
    payload = bytes(range(128)) if rank == 0 else None
    got = D.bcast_bytes(payload, 128)
    ok = got == bytes(range(128))
    # 2. reductions used for max-over-ranks timing and launch counts
    ok = ok and D.max(float(rank + 1)) == float(world) and D.sum(float(rank + 1)) == world * (world + 1) / 2
    D.barrier()
    # 3. my slab == my block of the reference decomposition
    X, B = bench.manufactured_slab(npglobal, world, rank)
    o = po.Oracle(po.make_config(npglobal, (1, 1, world), bcs=(0,) * 6))
    o.set_problem()
    ok = ok and np.abs(X - o.x(rank)).max() <= 1e-13 and np.abs(B - o.b(rank)).max() <= 1e-13
    bi = o.block(rank)
    ok = ok and list(bi.loc) == [0, 0, rank] and X.shape == o.shape(rank)
    # 4. gather of per-rank results on rank 0 (what mg_check.py does with the solutions)
    mine = torch.from_numpy(np.ascontiguousarray(X))
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank == 0:
        for r in range(world):
            ok = ok and np.abs(parts[r].numpy() - o.x(r)).max() <= 1e-13
    o.close()
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_two_ranks(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), (12, 10, 16), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_host_rank_layout():
    import bench
    assert bench.host_rank_layout((512, 512, 512), 16) == (1, 1, 16)
    assert bench.host_rank_layout((512, 512, 512), 8) == (1, 1, 8)
    assert bench.host_rank_layout((512, 512, 512), 1) == (1, 1, 1)
    assert bench.host_rank_layout((1024, 1024, 1024), 200) == (1, 1, 64)
    assert bench.parse_workload("auto", 1) == (512, 512, 512) and bench.parse_workload("auto", 8) == (1024, 1024, 1024)
    assert bench.parse_workload("96", 1) == (96, 96, 96) and bench.parse_workload("64x32x16", 2) == (64, 32, 16)
