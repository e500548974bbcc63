"""Multi-GPU plumbing: contiguous column blocks per rank, no collective on the timestep path, one
gather of per-column summaries at the end of a run.  Pure torch.distributed
(NCCL on the GPU box, gloo in the CPU tests) -- there is nothing to exchange while stepping because
columns share nothing (the reference keeps one model_state per object, include/all.hxx:264-275)."""
from __future__ import annotations


def shard_range(n_columns_total: int, rank: int, world: int):
    """GPU g owns columns [g*ceil(N/G), min(N, (g+1)*ceil(N/G))).  Returns (col0, ncol)."""
    per = -(-n_columns_total // world)
    col0 = min(n_columns_total, rank * per)
    return col0, max(0, min(n_columns_total, col0 + per) - col0)


def gather_summaries(local, dist=None, dst: int = 0):
    """Gather a per-column summary tensor [ncol_local, k] from every rank onto `dst` (in column
    order).  Ranks may own different column counts (last shard ragged).  Returns the concatenated
    tensor on dst and None elsewhere; without a process group returns `local`."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    padded = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def reduce_totals(local_totals, dist=None):
    """Sum of global mass-balance totals over ranks (all ranks get the result)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_totals
    out = local_totals.clone()
    dist.all_reduce(out)
    return out
