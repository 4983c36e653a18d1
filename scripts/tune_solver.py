"""Times one modal solve of the bench workload for a few solver settings (tuning aid, GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200.diffelastic.diff_model import DiffSoundObj
from diffsound_b200.diffelastic.deform import Deform

N = 32
dev = torch.device("cuda:0")
MESH = os.environ.get("MESH", "")          # "" = the bench cube; "bowl" / "grid16": fixtures of tests/golden/meshes.npz
MODES = int(os.environ.get("MODES", "32"))
if MESH:
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes.npz"))
    v, t = d[f"{MESH}_verts"], d[f"{MESH}_tets"].astype(np.int64)
else:
    v, t = bench.kuhn_cube(N)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=MODES, order=2, mat=bench.STEEL)
leaf = obj.tetmesh.vertices.detach().clone().requires_grad_(True)
obj.tetmesh.vertices = leaf


def solve():
    obj.deform = Deform(obj.tetmesh); obj._X = None; obj._warm = []; obj._Kval = obj._Mblk = None
    obj.eigen_decomposition()
    vals = obj.get_vals()
    leaf.grad = None
    (vals[:, 0] * (1.0 / obj.eigenvalues).float()).sum().backward()


for cfg in sys.argv[1:]:
    for kv in cfg.split(","):
        k, val = kv.split("=")
        setattr(DiffSoundObj, k, type(getattr(DiffSoundObj, k))(float(val)) if not isinstance(getattr(DiffSoundObj, k), bool) else bool(int(val)))
    for _ in range(2):
        solve()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    s = obj.eig_stats
    print(f"{cfg}: {dt * 1e3:.1f} ms/solve  fine its {s['iterations']}  nested its {s['nested_iterations']}  lam0 {float(obj.eigenvalues[0]):.8e}", flush=True)
