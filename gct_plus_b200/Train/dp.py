"""Host-side data-parallel helpers (one process per GPU, torch.distributed for the plumbing).

Training follows train1.py:111-112 (DDP): every rank computes the gradient of ITS sum-loss, gradients are
summed over ranks and divided by the world size (the division is folded into the fused Adam kernel).
Sampling shards independent latent draws with no data-path collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous share [lo, hi) of n_total independent work items for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    """Distinct, reproducible RNG stream per rank for the latent draws."""
    return (int(base_seed) * 1000003 + 7919 * int(rank)) % (2 ** 31 - 1)


def allreduce_sum_(flat: torch.Tensor, group=None) -> float:
    """In-place sum over ranks of the flat gradient buffer; returns the scale (1/world) that turns it into DDP's mean."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world
