"""Batch (view) sharding of the DIB-R path across the GPUs of one box.

Every view is independent in all four kernels (rasterization_cuda.cu:71-72,
dibr_soft_mask_cuda.cu:83), so the forward needs no
communication: rank r renders views [start, stop).  The only exchange is an
all-gather of the per-view gradients after backward
(``torch.distributed.all_gather_into_tensor`` — NCCL over NVLink on GPUs, gloo in
the CPU tests).  One process per GPU; nothing here touches the kernels.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_views", "all_gather_view_grads"]


def shard_range(batch, rank, world_size):
    """Contiguous shard [start, stop) of `batch` views for `rank`; the first
    ``batch % world_size`` ranks get one extra view."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(batch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_views(tensors, rank, world_size):
    """Slices the leading (view) dimension of every tensor for this rank."""
    out = []
    for t in tensors:
        s, e = shard_range(t.shape[0], rank, world_size)
        out.append(t[s:e])
    return out


def all_gather_view_grads(local_grads, batch, group=None):
    """All-gathers per-view gradient tensors (leading dim = local views) into
    full-batch tensors on every rank.  Equal shards use one
    ``all_gather_into_tensor`` per tensor; ragged shards are padded to the
    largest shard and trimmed."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    outs = []
    for g in local_grads:
        if g.shape[0] != sizes[rank]:
            raise ValueError(f"local gradient has {g.shape[0]} views, shard owns {sizes[rank]}")
        g = g.contiguous()
        if len(set(sizes)) == 1:
            full = torch.empty((batch,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
            dist.all_gather_into_tensor(full, g, group=group)
        else:
            m = max(sizes)
            pad = torch.zeros((m,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
            pad[:g.shape[0]] = g
            buf = torch.empty((world * m,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
            dist.all_gather_into_tensor(buf, pad, group=group)
            full = torch.cat([buf[r * m:r * m + sizes[r]] for r in range(world)], 0)
        outs.append(full)
    return outs
