"""One warm-up + one modal solve of the bench workload (for ncu; numbers printed here are not bench values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200.diffelastic.diff_model import DiffSoundObj
from diffsound_b200.diffelastic.deform import Deform

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
v, t = bench.kuhn_cube(N)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=2, mat=bench.STEEL)
leaf = obj.tetmesh.vertices.detach().clone().requires_grad_(True)
obj.tetmesh.vertices = leaf
for i in range(reps):
    obj.deform = Deform(obj.tetmesh); obj._X = None; obj._warm = []; obj._Kval = obj._Mblk = None
    torch.cuda.synchronize()
    print("MARK solve", i, flush=True)
    obj.eigen_decomposition()
    vals = obj.get_vals()
    leaf.grad = None
    (vals[:, 0] * (1.0 / obj.eigenvalues).float()).sum().backward()
    torch.cuda.synchronize()
print(obj.eig_stats)
