"""Host-side logic of the multi-GPU paths on CPU, world_size 2, gloo:
candidate sharding + ordered gather of the sweep path, slab bounds / owner map / packed column ids of
the row-partitioned path, and the 64-byte handle exchange the row partition performs at start-up."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsound_b200.parallel import gather_ordered, owner_of, shard_indices, slab_bounds, sweep_modal_solves
from diffsound_b200.parallel.rowpart import packed_column_map


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # sweep: 7 candidates, each "solve" returns (candidate id, fake eigenvalues)
        cands = [10.0 * i for i in range(7)]
        res = sweep_modal_solves(cands, lambda c: (c, [c + 1.0, c + 2.0]))
        assert [r[0] for r in res] == cands
        assert all(r[1] == [c + 1.0, c + 2.0] for r, c in zip(res, cands))
        # a rank that reports the wrong number of results is an error, not a silent mis-ordering
        try:
            gather_ordered([1], 7)
            ok = False
        except ValueError:
            ok = True
        # handle exchange of the row partition: every rank ends with every rank's 64 bytes
        mine = bytes([rank] * 64)
        got = [None] * world
        dist.all_gather_object(got, mine)
        assert got == [bytes([r] * 64) for r in range(world)]
        # slab ownership agrees on all ranks and tiles the node range
        b = slab_bounds(1001, world)
        t = torch.tensor(b)
        dist.broadcast(t, 0)
        assert t.tolist() == b
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sweep_and_rowpart_host_logic_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, True), (1, True)]


def test_shard_indices_cover_everything():
    for n in (0, 1, 5, 64):
        for world in (1, 2, 3, 8):
            allidx = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert allidx == list(range(n))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
    assert gather_ordered([3, 4], 2) == [3, 4]          # no process group: identity


@pytest.mark.parametrize("n,world", [(10, 3), (274625, 8), (8, 8), (1001, 2)])
def test_slab_bounds_and_packed_columns(n, world):
    b = slab_bounds(n, world)
    assert b[0] == 0 and b[-1] == n and len(b) == world + 1
    sizes = np.diff(b)
    assert sizes.min() >= 1 and sizes.max() - sizes.min() <= 1
    j = torch.arange(n)
    own = owner_of(j, b)
    assert bool(((torch.tensor(b)[own] <= j) & (j < torch.tensor(b)[own + 1])).all())
    packed = packed_column_map(n, b, "cpu").to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(packed >> 28, own)
    assert torch.equal(packed & 0x0FFFFFFF, j - torch.tensor(b)[own])
    with pytest.raises(ValueError):
        slab_bounds(3, 4)
    with pytest.raises(ValueError):
        slab_bounds(100, 9)
