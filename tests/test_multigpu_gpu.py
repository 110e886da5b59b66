"""N>1 on real GPUs: one process per GPU over NCCL, images sharded with no collective on the data path;
the optional all_gather of the (small) outputs must reproduce the single-GPU result BITWISE.
Skipped unless at least 2 GPUs are visible (run with `gpurun --gpus 2`)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import interpolate_antialiasing_b200 as aa
    from interpolate_antialiasing_b200 import sharding
    g = torch.Generator().manual_seed(0)
    x = torch.rand((5, 3, 270, 480), generator=g) * 255           # uneven shards: 3 + 2 images
    xl = sharding.local_slice(x, rank, world).cuda(rank).contiguous(memory_format=torch.channels_last)
    yl = aa.linear_forward(xl, (56, 56), False)                    # no collective here
    y = sharding.gather_outputs(yl.contiguous(), x.shape[0])       # optional, NCCL all_gather
    ok = True
    if rank == 0:
        whole = aa.linear_forward(x.cuda(0).contiguous(memory_format=torch.channels_last), (56, 56), False)
        ok = torch.equal(y, whole.contiguous())
    dist.barrier()
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


def test_sharded_forward_nccl_bitwise_equal():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
