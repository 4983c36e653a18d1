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

from diffsound_b200.parallel import (gather_ordered, owner_of, shard_indices, slab_bounds, sweep_modal_solves,
                                     sweep_modal_solves_dynamic, WorkQueue, gather_indexed)
from diffsound_b200.parallel.rowpart import packed_column_map
from diffsound_b200.parallel.synth import ShardedModalSynth, batch_slice


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
        # dynamic assignment (work queue in the group's store): every candidate exactly once whatever the pace of the ranks
        import time

        def slow(c):
            time.sleep(0.02 if (int(c) // 10 + rank) % 2 else 0.001)
            return (c, rank)
        res = sweep_modal_solves_dynamic(cands, slow)
        assert [r[0] for r in res] == cands and {r[1] for r in res} <= {0, 1}
        mine_idx = list(WorkQueue(5))              # a second queue (same construction order on both ranks) starts at zero
        both = [None] * world
        dist.all_gather_object(both, mine_idx)
        assert sorted(i for part in both for i in part) == list(range(5))
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
    assert list(WorkQueue(3)) == [0, 1, 2]
    assert gather_indexed([(1, "b"), (0, "a")], 2) == ["a", "b"]
    with pytest.raises(ValueError):
        gather_indexed([(0, "a"), (0, "b")], 2)
    assert sweep_modal_solves_dynamic([5, 6], lambda c: c * 2) == [10, 12]


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


def _cpu_render(amp, damp, freq, T, sr):
    """fp64 closed form of oscillator.py:297-304 on the CPU: a stand-in for ds_modal_synth_fwd in the plumbing test."""
    tau = (torch.arange(T, dtype=torch.float64) + 1.0) / sr
    basis = torch.exp(-damp.double()[:, None] * tau) * torch.sin(2 * np.pi * freq.double()[:, None] * tau)
    return (amp.double() @ basis).float()


def _cpu_render_bwd(amp, damp, freq, gy, sr):
    with torch.enable_grad():          # called from inside an autograd backward, where grad mode is off
        a, d, f = (t.double().clone().requires_grad_(True) for t in (amp, damp, freq))
        T = gy.shape[1]
        tau = (torch.arange(T, dtype=torch.float64) + 1.0) / sr
        y = a @ (torch.exp(-d[:, None] * tau) * torch.sin(2 * np.pi * f[:, None] * tau))
        y.backward(gy.double())
    return a.grad.float(), d.grad.float(), f.grad.float()


def _synth_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        B, k, T, sr = 5, 6, 300, 8000.0
        amp = torch.rand(B, k, generator=g)
        damp = (torch.rand(k, generator=g) * 20 + 1).requires_grad_(True)
        freq = (torch.rand(k, generator=g) * 2000 + 100).requires_grad_(True)
        sl = batch_slice(B, rank, world)
        a_loc = amp[sl].clone().requires_grad_(True)
        y = ShardedModalSynth.apply(a_loc, damp, freq, T, sr, _cpu_render, _cpu_render_bwd)
        (0.5 * (y ** 2).sum()).backward()
        # single-process reference over the full batch
        a_all = amp.clone().requires_grad_(True)
        d2, f2 = damp.detach().clone().requires_grad_(True), freq.detach().clone().requires_grad_(True)
        y_all = ShardedModalSynth.apply(a_all, d2, f2, T, sr, _cpu_render, lambda *a: _cpu_render_bwd(*a))
        ok = torch.allclose(y, y_all[sl], rtol=1e-6, atol=1e-7)
        # the reference run must not all-reduce twice: compute its gradient by hand
        ga, gd, gf = _cpu_render_bwd(amp, damp.detach(), freq.detach(), y_all.detach(), sr)
        ok = ok and torch.allclose(a_loc.grad, ga[sl], rtol=1e-5, atol=1e-6)
        ok = ok and torch.allclose(damp.grad, gd, rtol=1e-4, atol=1e-4 * float(gd.abs().max()))
        ok = ok and torch.allclose(freq.grad, gf, rtol=1e-4, atol=1e-4 * float(gf.abs().max()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_synthesis_plumbing_world2():
    """batch-sharded synthesis: local audio = slice of the full render; d/d(damp), d/d(freq) summed."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_synth_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == [(0, True), (1, True)]


def test_batch_slice_tiles_the_batch():
    for B in (1, 5, 1024):
        for world in (1, 2, 3, 8):
            rows = [i for r in range(world) for i in range(B)[batch_slice(B, r, world)]]
            assert rows == list(range(B))


def test_slab_solver_morton_numbering_is_a_symmetric_permutation():
    """The private Morton numbering of the slab solver (RowPartLOBPCG._setup_morton): the permuted pattern and the gather
    indices of the K / M values describe P K P^T, P M P^T for the node permutation P, and perm3 / inv3 carry iterate blocks
    there and back.  Pure index arithmetic: checked on the CPU against dense matrices."""
    import types
    from diffsound_b200.parallel.rowpart_lobpcg import RowPartLOBPCG
    rng = np.random.default_rng(0)
    n = 23
    coords = torch.tensor(rng.random((n, 3)), dtype=torch.float32)
    # a random symmetric block pattern with full diagonal; rows of different lengths
    adj = rng.random((n, n)) < 0.2
    adj = adj | adj.T | np.eye(n, dtype=bool)
    brow = np.concatenate([[0], np.cumsum(adj.sum(1))]).astype(np.int32)
    bcol = np.concatenate([np.nonzero(adj[i])[0] for i in range(n)]).astype(np.int32)
    nnzb = int(brow[-1])
    Kval = rng.standard_normal(9 * nnzb)
    Mblk = rng.standard_normal(nnzb)

    def dense(brow, bcol, Kval, Mblk):
        K, M = np.zeros((3 * n, 3 * n)), np.zeros((3 * n, 3 * n))
        for i in range(n):
            b0, deg = int(brow[i]), int(brow[i + 1] - brow[i])
            for p in range(deg):
                j = int(bcol[b0 + p])
                for c in range(3):
                    for d in range(3):
                        K[3 * i + c, 3 * j + d] = Kval[9 * b0 + c * 3 * deg + 3 * p + d]     # the reference's value order
                    M[3 * i + c, 3 * j + c] = Mblk[b0 + p]
        return K, M

    pat = types.SimpleNamespace(n_nodes=n, brow=torch.tensor(brow), bcol=torch.tensor(bcol))
    me = types.SimpleNamespace()
    pp = RowPartLOBPCG._setup_morton(me, pat, coords)
    perm = me.perm.numpy()
    assert sorted(perm.tolist()) == list(range(n)) and np.array_equal(me.inv.numpy()[perm], np.arange(n))
    Kp, Mp = dense(pp.brow.numpy(), pp.bcol.numpy(), Kval[me._kidx.numpy()], Mblk[me._midx.numpy()])
    K, M = dense(brow, bcol, Kval, Mblk)
    p3 = me._perm3.numpy()
    assert np.array_equal(Kp, K[np.ix_(p3, p3)]) and np.array_equal(Mp, M[np.ix_(p3, p3)])
    X = rng.standard_normal((3 * n, 4))
    assert np.array_equal(X[p3][me._inv3.numpy()], X)
    assert pp.nnzb == nnzb and int(pp.brow[-1]) == nnzb


@pytest.mark.parametrize("w,world", [(48, 2), (48, 4), (48, 8), (32, 3), (32, 8), (48, 5)])
def test_column_sharded_coarse_solve_slices_and_merges(w, world):
    """The slab solver splits the columns of the (replicated) P1 coarse solve over the ranks: every column is solved by
    exactly one rank, padding columns are zero, and the merge restores the column order."""
    from diffsound_b200.parallel.rowpart_lobpcg import coarse_column_merge, coarse_column_slice
    rc = torch.arange(7 * 64, dtype=torch.float32).reshape(7, 64) + 1.0
    parts = [coarse_column_slice(rc, w, r, world) for r in range(world)]
    assert all(p.shape == parts[0].shape and p.shape[1] % 16 == 0 for p in parts)
    per = -(-w // world)
    for r, p in enumerate(parts):
        c0, c1 = min(r * per, w), min(r * per + per, w)
        assert torch.equal(p[:, :c1 - c0], rc[:, c0:c1]) and not p[:, c1 - c0:].any()
    merged = coarse_column_merge(torch.stack(parts), w)          # "solve" = identity
    assert torch.equal(merged, rc[:, :w])
