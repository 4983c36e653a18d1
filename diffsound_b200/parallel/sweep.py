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


class WorkQueue:
    """Dynamic assignment of candidates: every rank pulls the next index from one atomic counter (the process group's
    key-value store; `add` is a fetch-and-add over TCP, ~0.1 ms against ~1 s per modal solve).  For sweeps whose candidates
    differ widely in cost -- the marching-tets thickness sweep needs 23 .. 300 LOBPCG iterations per candidate, round-robin
    left the slowest of 8 ranks with 2.1 x the time of the fastest -- this keeps every rank busy until the list is empty.
    Every rank must construct its queues in the same order (the counter key carries a sequence number)."""
    _seq = 0

    def __init__(self, n_items: int):
        self.n = int(n_items)
        WorkQueue._seq += 1
        self._key = f"diffsound_b200/work_queue/{WorkQueue._seq}"
        self._store = None
        self._next = 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._store = dist.distributed_c10d._get_default_store()

    def __iter__(self):
        while True:
            if self._store is None:
                i = self._next
                self._next += 1
            else:
                i = int(self._store.add(self._key, 1)) - 1
            if i >= self.n:
                return
            yield i


def gather_indexed(local_pairs: Sequence, n_items: int) -> List:
    """All ranks' results in candidate order from (index, result) pairs in any assignment; every index exactly once."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, list(local_pairs))
    else:
        parts = [list(local_pairs)]
    out = [None] * n_items
    seen = 0
    for part in parts:
        for idx, res in part:
            if not (0 <= idx < n_items) or out[idx] is not None:
                raise ValueError(f"candidate {idx} reported twice or out of range")
            out[idx] = res
            seen += 1
    if seen != n_items:
        raise ValueError(f"{seen} results for {n_items} candidates")
    return out


def sweep_modal_solves_dynamic(candidates: Sequence, solve: Callable) -> List:
    """Like sweep_modal_solves, but candidates are handed out on demand (WorkQueue) instead of round-robin."""
    pairs = [(i, solve(candidates[i])) for i in WorkQueue(len(candidates))]
    return gather_indexed(pairs, len(candidates))
