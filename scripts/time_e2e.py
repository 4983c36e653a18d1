"""Where the end-to-end time of one modal solve goes (host clock with a device sync after every stage;
diagnostic only -- bench.py's e2e number has no syncs inside)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffsound_b200.diffelastic.diff_model import DiffSoundObj
from diffsound_b200.diffelastic.mesh import TetMesh

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
v, t = bench.kuhn_cube(N)
vh, th = torch.from_numpy(v).pin_memory(), torch.from_numpy(t).pin_memory()


def tick(name, t0, acc):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3
    return t1


for rep in range(6):
    acc = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    leaf = vh.to(dev, non_blocking=True).requires_grad_(True)
    td = th.to(dev, non_blocking=True)
    t0 = tick("h2d", t0, acc)
    m2 = TetMesh(leaf, td).to_high_order(2)
    t0 = tick("promote (standalone)", t0, acc)
    obj = DiffSoundObj(leaf, td, mode_num=32, order=2, mat=bench.STEEL)
    t0 = tick("DiffSoundObj() incl. promote", t0, acc)
    _ = obj.deform.pattern
    t0 = tick("pattern", t0, acc)
    _ = obj.deform.incidence
    t0 = tick("incidence", t0, acc)
    _ = obj.deform.coarse
    t0 = tick("coarse level", t0, acc)
    obj.eigen_decomposition()
    t0 = tick("assemble + lobpcg", t0, acc)
    vals = obj.get_vals()
    (vals[:, 0] * (1.0 / obj.eigenvalues).float()).sum().backward()
    t0 = tick("get_vals + backward", t0, acc)
    lam_h = obj.eigenvalues.cpu()
    g_h = leaf.grad.cpu()
    t0 = tick("d2h", t0, acc)
    st = torch.cuda.memory_stats()
    acc["cudaMallocs"] = st.get("num_device_alloc", 0)
    acc["reserved GB"] = st.get("reserved_bytes.all.current", 0) / 1e9
    if rep:
        print("  ".join(f"{k}: {v:.2f} ms" for k, v in acc.items()), flush=True)
