"""Sharding of independent modal solves across ranks.

Reference behaviour replaced: the sequential candidate loops of the experiment scripts
(experiments/thickness_train.py:127-141, experiments/material_sync_train.py:95-118 of the reference): every
candidate builds its own mesh / model and is solved independently, so candidate i simply goes to rank
i mod world; the only communication is one gather of the per-candidate results at the end.
"""
from typing import Callable, List, Sequence

import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of the candidates rank `rank` solves (round-robin: neighbouring candidates of a sweep have
    similar mesh sizes, so round-robin balances the load better than contiguous chunks)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def gather_ordered(local: Sequence, n_items: int, rank: int = None, world: int = None) -> List:
    """All ranks' results in candidate order.  `local[k]` belongs to candidate shard_indices(...)[k]."""
    if not (dist.is_available() and dist.is_initialized()):
        if len(local) != n_items:
            raise ValueError("single process: local results must cover every candidate")
        return list(local)
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    mine = shard_indices(n_items, rank, world)
    if len(local) != len(mine):
        raise ValueError(f"rank {rank}: {len(local)} results for {len(mine)} candidates")
    parts = [None] * world
    dist.all_gather_object(parts, list(local))
    out = [None] * n_items
    for r, part in enumerate(parts):
        for k, idx in enumerate(shard_indices(n_items, r, world)):
            out[idx] = part[k]
    return out


def sweep_modal_solves(candidates: Sequence, solve: Callable, rank: int = None, world: int = None) -> List:
    """Run `solve(candidate)` for this rank's share of `candidates` and return all results in order on
    every rank.  `solve` must return something picklable and small (eigenvalues, scalar gradients);
    big tensors stay on the rank that produced them."""
    if dist.is_available() and dist.is_initialized():
        rank = dist.get_rank() if rank is None else rank
        world = dist.get_world_size() if world is None else world
    else:
        rank, world = 0, 1
    local = [solve(candidates[i]) for i in shard_indices(len(candidates), rank, world)]
    return gather_ordered(local, len(candidates), rank, world)
