"""Frame sharding for multi-GPU runs: frames and clips are independent, so
rank g of G owns the contiguous block [g*n/G, (g+1)*n/G) and no data-path collective exists.
torch.distributed is used only for the barrier and for the max-over-ranks of a timing."""


def frame_shard(n_frames, rank, world):
    """Contiguous block of frames owned by `rank`; blocks differ by at most one frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('bad rank/world: %r/%r' % (rank, world))
    base, extra = divmod(int(n_frames), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
