"""Row-partitioned FP32 SpMM (PEER variant of k_spmm32v) against the single-GPU kernel: identical
accumulation order per row, so the slabs must be BIT-EXACT.  world = 1 runs in the normal GPU suite;
world = 2 needs two GPUs (gpurun --gpus 2) and is skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
STEEL = (7850.0, 2.0e11, 0.29, 20, 3e-8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(dev, N=6):
    import bench
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj
    v, t = bench.kuhn_cube(N)
    obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=8, order=2, mat=STEEL)
    obj._assemble(obj.material_model.mat.density)
    return obj


def _check(rank, world, dev, ncols):
    import torch.distributed as dist
    from diffsound_b200 import native
    from diffsound_b200.parallel import RowPartition
    obj = _problem(dev)
    pat = obj.deform.pattern
    g = torch.Generator(device=dev).manual_seed(5)
    X = torch.randn(pat.n, ncols, device=dev, generator=g)
    R = torch.randn(pat.n, ncols, device=dev, generator=g) * 1e9
    Zp = torch.randn(pat.n, ncols, device=dev, generator=g)
    rec, invD = native.k32_pack(pat, obj._Kval)
    part = RowPartition(pat, obj._Kval, ncols, nbuf=2)
    lo, hi = 3 * part.bounds[rank], 3 * part.bounds[rank + 1]
    part.blocks[0].copy_(X[lo:hi])
    part.blocks[1].copy_(Zp[lo:hi])
    part.barrier()
    # plain product into a local tensor
    y = part.spmm(0, torch.empty(hi - lo, ncols, device=dev), mode=0)
    ref = native.spmm32(pat, rec, X, mode=0)
    assert torch.equal(y, ref[lo:hi]), float((y - ref[lo:hi]).abs().max())
    assert np.allclose(part.invD.cpu().numpy(), invD.cpu().numpy()[9 * part.bounds[rank]:9 * part.bounds[rank + 1]])
    # one Chebyshev step, output aliasing Zprev in the second peer block
    ref2 = Zp.clone()
    native.spmm32(pat, rec, X, mode=2, R=R, invD=invD, Zprev=ref2, ab=0.4, cc=1e-10, out=ref2)
    part.spmm(0, part.blocks[1], mode=2, R=R[lo:hi].contiguous(), Zprev=part.blocks[1], ab=0.4, cc=1e-10)
    assert torch.equal(part.blocks[1], ref2[lo:hi])
    # the freshly written block is the gather source of the next product on every rank
    part.barrier()
    y2 = part.spmm(1, torch.empty(hi - lo, ncols, device=dev), mode=0)
    ref3 = native.spmm32(pat, rec, ref2, mode=0)
    assert torch.equal(y2, ref3[lo:hi])
    torch.cuda.synchronize(dev)
    part.close()


def _worker(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        for ncols in (16, 48):
            _check(rank, world, dev, ncols)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncols", [16, 32, 48])
def test_rowpart_world1_bit_exact(ncols):
    _check(0, 1, torch.device("cuda:0"), ncols)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_rowpart_world2_bit_exact():
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


# ---- LOBPCG on row slabs (parallel/rowpart_lobpcg.py): the same eigenpairs as the single-GPU driver ---------------------
def _lobpcg_case(dev, group=None, N=6, k=10):
    import bench
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj
    from diffsound_b200.parallel.rowpart_lobpcg import eigen_decomposition_rowpart
    v, t = bench.kuhn_cube(N)
    vd, td = torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev)
    ref = DiffSoundObj(vd, td, mode_num=k, order=2, mat=STEEL)
    ref.eig_tol = 1e-7
    ref.eigen_decomposition()
    obj = DiffSoundObj(vd, td, mode_num=k, order=2, mat=STEEL)
    obj.eig_tol = 1e-7
    stats = eigen_decomposition_rowpart(obj, group=group)
    lam_ref, lam = ref.eigenvalues.cpu().numpy(), obj.eigenvalues.cpu().numpy()
    # residual 1e-7 -> eigenvalue error ~1e-14 relative to the spectrum; both solves sit on the same pairs
    assert np.abs(lam - lam_ref).max() / lam_ref.max() <= 1e-10, (lam, lam_ref)
    # M-orthonormal block of full height on every rank, and a usable backward
    U = obj.U_hat_full
    assert U.shape == ref.U_hat_full.shape
    MU = torch.empty_like(U)
    KU = torch.empty_like(U)
    from diffsound_b200 import native
    native.spmm_k_and_m(obj.deform.pattern, obj._Kval, obj._Mblk, U.contiguous(), KU, MU)
    G = (U.T @ MU).cpu().numpy()
    assert np.abs(G - np.eye(G.shape[0])).max() <= 1e-10
    res = (KU[:, 6:] - MU[:, 6:] * obj.eigenvalues[None, :]).norm(dim=0) / (obj.eigenvalues * MU[:, 6:].norm(dim=0))
    assert float(res.max()) <= 1e-6
    return stats


def test_rowpart_lobpcg_world1_matches_single_gpu_driver():
    stats = _lobpcg_case(torch.device("cuda:0"))
    assert stats["status"] == 0 and stats["world"] == 1


def _lobpcg_worker(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        stats = _lobpcg_case(dev)
        assert stats["status"] == 0 and stats["world"] == world
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_rowpart_lobpcg_world2_matches_single_gpu_driver():
    import torch.multiprocessing as mp
    mp.spawn(_lobpcg_worker, args=(2, _free_port()), nprocs=2, join=True)
