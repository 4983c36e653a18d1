"""Thickness sweep of BASELINE.json configs[3]: 64 hollow-sphere candidates on the 64^3 Quartet
background grid (x 1.5), shell 0 < sdf <= t max(sdf), t = linspace(0.2, 0.9, 64); per candidate: marching tets -> largest
component -> quadratic promotion -> pattern + assembly -> LOBPCG (32 modes + 6 rigid) -> get_vals -> backward to the
thickness coefficient.  Candidates are independent modal solves: rank r takes candidates r, r + world, ... through
diffsound_b200.parallel.sweep (no data-path collective; NCCL only for the barrier, the max-over-ranks time and the final
gather of 32 eigenvalues + one gradient per candidate).  Prints one JSON line (rank 0).

usage: python scripts/bench_sweep.py [n_candidates] [reps] [dynamic|roundrobin]   (assignment of candidates to ranks;
       default dynamic: a shared counter hands out the next candidate, parallel/sweep.py WorkQueue)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_sweep.py"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from diffsound_b200.dmtet.geometry.dmtet_thickness import DMTetGeometry
from diffsound_b200.parallel.sweep import WorkQueue, gather_indexed, gather_ordered, shard_indices

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
assign = sys.argv[3] if len(sys.argv) > 3 else "dynamic"
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)

d = np.load(os.path.join(ROOT, "tests", "golden", "grid64_tets.npz"))
FLAGS = types.SimpleNamespace(mode_num=32, order=2, mat="Steel", out_dir="/tmp", without_tensorboard=True)
geo = DMTetGeometry(64, 1.5, FLAGS, grid=(d["vertices"], d["indices"]))
geo.apply_sdf(lambda v: 0.6 - v.norm(dim=1))           # sphere, positive inside (dmtet_thickness.py:312)
coefs = torch.linspace(0.2, 0.9, n_cand).tolist()
mine = shard_indices(n_cand, rank, world)


def solve(coef):
    tc = torch.tensor(coef, requires_grad=True)
    obj = geo.getMesh(thickness_coef=tc)
    obj.eigen_decomposition()
    vals = obj.get_vals()
    ((vals[:, 0] / obj.eigenvalues.float() - 1.0) ** 2).mean().backward()
    return obj, tc


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for c in (coefs[mine[0]], coefs[mine[-1]]):            # warm-up: allocator + workspace sized for the largest mesh
    solve(c)
sizes, results = [], []
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    results, sizes, taken = [], [], []
    for i in (WorkQueue(n_cand) if assign == "dynamic" else mine):
        obj, tc = solve(coefs[i])
        taken.append(i)
        results.append((obj.eigenvalues.cpu().tolist(), float(tc.grad)))
        sizes.append((int(obj.tetmesh.tets.shape[0]), int(obj.deform.pattern.n), int(obj.eig_stats["iterations"])))
e1.record()
torch.cuda.synchronize()
ms_local = e0.elapsed_time(e1) / reps
t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
tmin = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
all_res = gather_indexed(list(zip(taken, results)), n_cand)
all_sizes = gather_indexed(list(zip(taken, sizes)), n_cand)
if rank == 0:
    lam0 = [r[0][0] for r in all_res]
    line = {"what": "thickness sweep (BASELINE configs[3])", "n_gpus": world, "candidates": n_cand, "scaling": "strong (candidates sharded)", "assignment": assign,
            "value": n_cand / (float(t.item()) * 1e-3), "unit": "solves/s", "ms_sweep_max_rank": float(t.item()),
            "ms_sweep_min_rank": float(tmin.item()), "load_imbalance_max_over_min": float(t.item()) / float(tmin.item()),
            "tets_min_max": [min(s[0] for s in all_sizes), max(s[0] for s in all_sizes)],
            "dofs_min_max": [min(s[1] for s in all_sizes), max(s[1] for s in all_sizes)],
            "lobpcg_iterations_min_max": [min(s[2] for s in all_sizes), max(s[2] for s in all_sizes)],
            "first_eigenvalue_first_last": [lam0[0], lam0[-1]], "dgrad_first_last": [all_res[0][1], all_res[-1][1]],
            "config": "64^3 Quartet grid x 1.5, sphere SDF r = 0.6, t = linspace(0.2, 0.9, %d), order 2, 32 modes, Steel" % n_cand}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
