"""Micro-benchmark of the fine-level FP32 SpMM (k_spmm32, Chebyshev mode) on the bench mesh:
CUDA-event time per launch and achieved algorithmic GB/s.  Also the target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200 import native
from diffsound_b200.diffelastic.diff_model import DiffSoundObj

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
v, t = bench.kuhn_cube(N)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=2, mat=bench.STEEL)
obj._assemble(obj.material_model.mat.density)
pat = obj.deform.pattern
rec, invD = native.k32_pack(pat, obj._Kval)
for c in (64, 48, 32, 16):
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
    print(f"c={c}: {ms * 1e3:.1f} us/launch  {nbytes / ms / 1e6:.0f} GB/s algorithmic ({nbytes / 1e6:.0f} MB)", flush=True)
