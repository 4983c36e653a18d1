"""Micro-benchmark of the fine-level FP32 SpMM (Chebyshev mode) on the bench mesh: CUDA-event time per launch
and achieved algorithmic GB/s, for the reference (lexicographic) and a Morton node numbering.  Also the target of the ncu captures.

usage: python scripts/bench_spmm32.py [cube N] [reps] [orders e.g. lm] [cols e.g. 48,16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200 import native
from diffsound_b200.diffelastic.diff_model import DiffSoundObj
from diffsound_b200.diffelastic.deform import Deform

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
orders = sys.argv[3] if len(sys.argv) > 3 else "lm"
cols = [int(c) for c in (sys.argv[4] if len(sys.argv) > 4 else "48,32,16").split(",")]
dev = torch.device("cuda:0")
lib = native._lib.load()


def morton_perm(v):
    q = ((v - v.min(0).values) / (v.max(0).values - v.min(0).values + 1e-30) * 1023).long()
    code = torch.zeros(v.shape[0], dtype=torch.long, device=v.device)
    for b in range(10):
        for d in range(3):
            code |= ((q[:, d] >> b) & 1) << (3 * b + (2 - d))
    return torch.argsort(code, stable=True)


v, t = bench.kuhn_cube(N)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=2, mat=bench.STEEL)
for order in orders:
    if order == "m":       # renumber the promoted mesh along a Morton curve and rebuild pattern + matrices
        perm = morton_perm(obj.tetmesh.vertices.detach())
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=dev)
        obj.tetmesh.vertices = obj.tetmesh.vertices.detach()[perm].contiguous()
        obj.tetmesh.tets = inv[obj.tetmesh.tets].contiguous()
        obj.deform = Deform(obj.tetmesh)
        obj._Kval = obj._Mblk = None
    obj._assemble(obj.material_model.mat.density)
    pat = obj.deform.pattern
    rec, invD = native.k32_pack(pat, obj._Kval)
    for c in cols:
        X = torch.randn(pat.n, c, device=dev)
        R = torch.randn(pat.n, c, device=dev)
        Zp = torch.randn(pat.n, c, device=dev)
        out = torch.empty_like(X)
        for _ in range(3):
            native.spmm32(pat, rec, X, mode=2, R=R, invD=invD, Zprev=Zp, ab=0.3, cc=1e-12, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            native.spmm32(pat, rec, X, mode=2, R=R, invD=invD, Zprev=Zp, ab=0.3, cc=1e-12, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = pat.nnzb * 40 + pat.n_nodes * 40 + 4 * pat.n * c * 4
        print(f"order={order} c={c}: {ms * 1e3:.1f} us/launch  {nbytes / ms / 1e6:.0f} GB/s algorithmic "
              f"({nbytes / 1e6:.0f} MB)", flush=True)
