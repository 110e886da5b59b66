"""world_size-2 gloo test of the N>1 path's host logic: shard by image, apply, gather == whole batch.
The op itself is stood in for by the oracle here (no GPU in the build container); on the GPU box the
same helpers wrap aa.linear_forward (tests/test_multigpu_gpu.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from interpolate_antialiasing_b200 import sharding
    from oracle import aa_oracle as O
    g = torch.Generator().manual_seed(0)
    x = torch.rand((n, 3, 20, 24), generator=g) * 255
    op = lambda t: torch.from_numpy(O.forward(t.numpy(), (7, 9), "linear", False)) if t.shape[0] else torch.empty((0, 3, 7, 9))
    y = sharding.sharded_apply(x, op, gather=True)
    whole = op(x)
    ok = torch.equal(y, whole)  # bitwise: shards are independent
    b, e = sharding.shard_bounds(n, world)[rank]
    ok = ok and torch.equal(sharding.local_slice(x), x[b:e])
    dist.barrier()
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


def _run(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shard_bounds():
    from interpolate_antialiasing_b200.sharding import shard_bounds
    assert shard_bounds(256, 8) == [(32 * r, 32 * r + 32) for r in range(8)]
    assert shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    for n in range(0, 40):
        for w in range(1, 9):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_sharded_apply_even_gloo_world2():
    _run(6)


def test_sharded_apply_uneven_gloo_world2():
    _run(5)
