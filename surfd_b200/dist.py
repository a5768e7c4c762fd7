"""Multi-GPU plumbing of the generation path: independent shapes sharded over ranks, one start-up broadcast of the packed
weights, no data-path collective.  Mirrors what utils/dist_util.py:61-67 (`sync_params`) does per tensor,
as a single flat buffer per blob.  Works with backend 'nccl' (GPU box) and 'gloo' (CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, world, rank):
    """contiguous shard [lo, hi) of `n_items` for `rank` (rank r gets samples [r*ceil(n/world), ...))"""
    per = (n_items + world - 1) // world
    return min(n_items, rank * per), min(n_items, (rank + 1) * per)


def sliced_noise(seed, n_steps, batch_total, latent, lo, hi):
    """x_T + per-step randn_like draws from one CPU generator in full-batch order, then sliced: every GPU count (and the
    oracle) sees the same noise for the same sample."""
    g = torch.Generator().manual_seed(seed)
    full = torch.randn(n_steps + 1, batch_total, latent, generator=g)
    return full[:, lo:hi].contiguous()


def broadcast_packed(tensors, src=0):
    """in-place broadcast of a list of flat tensors (weights blob, op program, decoder blob) from rank `src`"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    for t in tensors:
        dist.broadcast(t, src)
    return tensors


def conditioning_from_rank0(make, batch, width, device, rank=None, world=None):
    """the [batch, width] conditioning tensor (CLIP embeddings): rank 0 alone runs `make()` -- one encoder pass per generation,
    one copy of the encoder weights per node -- and the result reaches the other ranks through one broadcast; a single
    process just calls `make()`."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    if world == 1:
        return make()
    if rank == 0:
        ctx = make().to(device, torch.float32).contiguous()
        if tuple(ctx.shape) != (batch, width):
            raise ValueError(f"conditioning must have shape ({batch}, {width}), got {tuple(ctx.shape)}")
    else:
        ctx = torch.empty(batch, width, device=device, dtype=torch.float32)
    dist.broadcast(ctx, src=0)
    return ctx
