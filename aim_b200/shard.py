"""One process per GPU: how pairs are split over ranks and how per-rank timings are combined.

Pairs are independent, so the path shards by pair index with NO data-path collective (the reference
does the same over DPUs: WFA/DPU-MRAM/host/host.c:191-209).  torch.distributed is used only for the
barrier and for max-over-ranks timing (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def rank_slice(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [first, first+count) of `total` pairs for `rank` (strong split, like host.c per DPU)."""
    per = -(-total // world)
    first = min(total, rank * per)
    return first, min(per, total - first)


def weak_first_pair(rank: int, pairs_per_rank: int) -> int:
    """Weak scaling: every rank aligns its own `pairs_per_rank` pairs of one global synthetic stream."""
    return rank * pairs_per_rank


def max_over_ranks(seconds: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_pair_order(local, world: int, rank: int):
    """Concatenate per-rank result arrays on rank 0 in pair order (test/report helper; the product
    writes each shard straight into the caller's arrays at its offset)."""
    import numpy as np
    import torch.distributed as dist
    if world == 1:
        return local
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0)
    return np.concatenate(out) if rank == 0 else None
