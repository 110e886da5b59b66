"""Batch sharding across the GPUs of one box: one process per GPU, images are independent, so the
hot path needs NO collective.  NCCL (or gloo in the CPU tests) is used only for the
optional gathering of the small outputs and for barriers/timing reductions."""
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous chunks of ceil(n/world) images: [(begin, end)] * world (trailing shards may be empty)."""
    per = -(-n // world) if world > 0 else n
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def local_slice(x, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b, e = shard_bounds(x.shape[0], world)[rank]
    return x[b:e]


def gather_outputs(y_local, n_total, group=None):
    """all_gather of per-rank outputs (uneven last shards are padded); returns the [n_total, ...] batch."""
    world = dist.get_world_size(group)
    bounds = shard_bounds(n_total, world)
    per = bounds[0][1] - bounds[0][0]
    pad = torch.zeros((per,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: e - b] for p, (b, e) in zip(parts, bounds)], dim=0)


def sharded_apply(x_full, op, gather=True, group=None):
    """Runs `op` on this rank's images of `x_full` (every rank holds or can index the batch) and
    optionally gathers.  `op` is e.g. lambda t: aa.linear_forward(t.cuda(), size, False)."""
    y = op(local_slice(x_full, dist.get_rank(group), dist.get_world_size(group)))
    return gather_outputs(y, x_full.shape[0], group) if gather else y
