"""The N > 1 path on CPU: world_size-2 gloo processes shard the shift grid / the z batch, 'compute' a stand-in for the
kernel, gather, and must reproduce the single-process answer and the max-over-ranks timing rule."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _standin(lam):
    # any deterministic per-shift function: the sharding logic is what is under test
    return np.sqrt(np.abs(lam) ** 2 + 4.0)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import ila_b200                                    # noqa: F401  (package import must work without a GPU)
    from ila_b200 import sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny = 16, 12
    lam = sh.grid_rows(nx, ny, rank, world)
    part = _standin(lam)
    rows = sh.interleaved(ny, rank, world)
    zidx = sh.interleaved(37, rank, world)
    cidx = sh.contiguous(37, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (part, rows, zidx, cidx))
    t = sh.max_over_ranks(1.0 + rank, dist)
    if rank == 0:
        q.put((gathered, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    from ila_b200 import sharding as sh
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29533 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nx, ny = 16, 12
    full = sh.grid_rows(nx, ny, 0, 1)
    ref = _standin(full).reshape(ny, nx)
    out = np.empty((ny, nx))
    for part, rows, _, _ in gathered:
        out[rows] = part.reshape(len(rows), nx)
    assert np.array_equal(out, ref)
    # every index exactly once, for both partitions
    for k in (2, 3):
        allidx = np.sort(np.concatenate([g[k] for g in gathered]))
        assert np.array_equal(allidx, np.arange(37))
    assert tmax == 2.0          # max over ranks of (1 + rank)
    parts = [g[0][:5] for g in gathered]
    idx = [np.arange(r, 10, 2) for r in range(2)]
    back = sh.scatter_back(parts, idx, 10)
    assert np.array_equal(back[0::2], parts[0]) and np.array_equal(back[1::2], parts[1])
