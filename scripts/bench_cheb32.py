"""Coarse-level Chebyshev solve (P1 operator of the bench mesh): one cooperative launch for all steps
(k_cheb32_persistent) against one k_spmm32v launch per step.  CUDA-event time per solve and per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200 import native
from diffsound_b200.diffelastic.diff_model import DiffSoundObj

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
v, t = bench.kuhn_cube(N)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=1, mat=bench.STEEL)
obj._assemble(obj.material_model.mat.density)
pat = obj.deform.pattern
rec, invD = native.k32_pack(pat, obj._Kval)
for c in (48, 32, 16):
    R = torch.randn(pat.n, c, device=dev)
    for persistent in (True, False):
        for _ in range(3):
            z = native.cheb32_solve(pat, rec, invD, R, degree, 2.5, 0.4 * degree * degree, persistent)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            z = native.cheb32_solve(pat, rec, invD, R, degree, 2.5, 0.4 * degree * degree, persistent)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"n_nodes={pat.n_nodes} nnzb={pat.nnzb} c={c} degree={degree} persistent={persistent}: {ms * 1e3:.0f} us/solve, "
              f"{ms * 1e3 / degree:.1f} us/step", flush=True)
