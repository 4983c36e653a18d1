"""Strong scaling of the row-partitioned FP32 SpMM (Chebyshev step, k_spmm32v PEER variant) on the
bench mesh (BASELINE configs[2]: 'row-partitioned SpMM at 1/2/4/8 GPUs').  Launch with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_rowpart.py
Every rank assembles the full matrix (replicated) and keeps only its slab of FP32 records; a step is
one Chebyshev step on all slabs + the cross-rank ordering barrier.  Prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from diffsound_b200.diffelastic.diff_model import DiffSoundObj
from diffsound_b200.parallel import RowPartition

rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("CUBE", "32"))
steps = int(os.environ.get("STEPS", "40"))
ncols = int(os.environ.get("NCOLS", "48"))
v, t = bench.kuhn_cube(N)
obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=2, mat=bench.STEEL)
obj._assemble(obj.material_model.mat.density)
pat = obj.deform.pattern
part = RowPartition(pat, obj._Kval, ncols, nbuf=2)
obj._Kval = None
R = torch.randn(3 * part.n_local, ncols, device=dev)
part.blocks[0].normal_()
part.blocks[1].normal_()


def run(k):
    for i in range(k):
        src, dst = i & 1, (i & 1) ^ 1
        part.barrier()
        part.spmm(src, part.blocks[dst], mode=2, R=R, Zprev=part.blocks[dst], ab=0.3, cc=1e-13)


run(4)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(steps)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
nbytes = pat.nnzb * 40 + pat.n_nodes * 40 + 4 * pat.n * ncols * 4
halo = 0
if rank == 0:
    print(json.dumps({"what": "row-partitioned k_spmm32v Chebyshev step, strong scaling", "n_gpus": world, "ncols": ncols,
                      "ms_per_step": float(ms), "algorithmic_GB_per_step": nbytes / 1e9,
                      "aggregate_GB_per_s": nbytes / float(ms) / 1e6, "n": pat.n, "nnz": 9 * pat.nnzb,
                      "slab_rows": [b - a for a, b in zip(part.bounds[:-1], part.bounds[1:])]}), flush=True)
part.close()
if world > 1:
    dist.destroy_process_group()
