"""Diagnostic: wall time per row-partitioned solve, host-side profile of one solve (cProfile), synchronised phase times."""
import os, sys, time, json, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200.diffelastic.diff_model import DiffSoundObj
from diffsound_b200.parallel.rowpart_lobpcg import eigen_decomposition_rowpart

dev = torch.device("cuda:0")
v, t = bench.kuhn_cube(32)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=2, mat=bench.STEEL)
solver = None
def solve():
    global solver
    obj._X = None; obj._warm = []; obj._Kval = obj._Mblk = None
    st, solver = eigen_decomposition_rowpart(obj, solver=solver, keep=True)
for _ in range(3):
    solve()
def timed(n=3):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); solve(); torch.cuda.synchronize()
        out.append(round((time.perf_counter() - t0) * 1e3, 1))
    return out
print("wall ms", timed(6))
pr = cProfile.Profile()
pr.enable(); solve(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4000])
solver.profile = True; solver.phase_ms = {}
solve(); print(json.dumps(solver.phase_ms))
solver.profile = False
solver.close()
