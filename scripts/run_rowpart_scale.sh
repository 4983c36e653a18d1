#!/bin/bash
# strong scaling of ONE 196 608-tet mesh: bench.py --mode rowpart at N = 1, 2, 4, 8 (as many as the box has)
tag=${1:-r2}
ng=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $ng ] && break
  if [ $n -eq 1 ]; then
    python bench.py --mode rowpart --steps 5 --warmup 3 > gpurun_out/${tag}_rowpart_n$n.json 2> gpurun_out/${tag}_rowpart_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --mode rowpart --steps 5 --warmup 3 > gpurun_out/${tag}_rowpart_n$n.json 2> gpurun_out/${tag}_rowpart_n$n.err
  fi
  tail -2 gpurun_out/${tag}_rowpart_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_rowpart_n$n.json").read().strip().splitlines()[-1])
    print($n, "ms/step", round(d["ms_per_step"],2), "dlam", d["config"]["max_rel_dlambda_vs_single_gpu_driver"], {k: round(v,1) for k,v in d["phase_wall_ms_rank0_synchronised"].items()})
except Exception as e:
    print($n, "failed", e)
PY
done
