"""Multi-GPU plumbing for the batched path: instances are independent, so the batch
is split into contiguous theta ranges, one per rank; every rank holds a replica of the KB-scale
constant block; the ONLY inter-GPU traffic is a final result gather.  torch.distributed is used
for the plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`:  [g*N/W, (g+1)*N/W)  (ragged sizes allowed)."""
    if world <= 0 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def gather_columns(local, n_total: int, dst: int = 0, group=None):
    """Final result gather: `local` is this rank's (rows x n_local) slice (torch tensor, any
    device); rank `dst` gets the (rows x n_total) result, the others None.  Ragged shards are
    padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    rows = local.shape[0]
    pad = torch.zeros((width, rows), dtype=local.dtype, device=local.device)
    pad[: local.shape[1]] = local.T          # instance-major so each shard is contiguous
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty((n_total, rows), dtype=local.dtype, device=local.device)
    for (lo, hi), b in zip(sizes, bufs):
        out[lo:hi] = b[: hi - lo]
    return out.T


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule for multi-GPU numbers: the slowest rank defines the step time."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
