#!/bin/bash
# Multi-GPU pass on one box (gpurun --gpus 8): sweep bench (weak scaling, the driver's SCALE configuration), ONE mesh on row
# slabs (strong scaling), the 64-candidate thickness sweep of BASELINE configs[3], batch-sharded synthesis (config 5).
# usage: bash scripts/gpu_multi.sh TAG "2 4 8"
TAG=${1:-r2m}
NS=${2:-"8"}
mkdir -p gpurun_out
ng=$(nvidia-smi -L | wc -l)
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) "$@"; }
for n in $NS; do
  [ $n -gt $ng ] && continue
  timeout 300 bash -c "$(declare -f run); run $n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline" > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err; echo "bench n=$n rc=$?"
  timeout 300 bash -c "$(declare -f run); run $n bench.py --gpus $n --mode rowpart --steps 5 --warmup 3" > gpurun_out/${TAG}_rowpart_n$n.json 2> gpurun_out/${TAG}_rowpart_n$n.err; echo "rowpart n=$n rc=$?"
  timeout 400 bash -c "$(declare -f run); run $n scripts/bench_sweep.py 64 1" > gpurun_out/${TAG}_sweep_n$n.json 2> gpurun_out/${TAG}_sweep_n$n.err; echo "sweep n=$n rc=$?"
  timeout 200 bash -c "$(declare -f run); run $n scripts/bench_synth.py" > gpurun_out/${TAG}_synth_n$n.json 2> gpurun_out/${TAG}_synth_n$n.err; echo "synth n=$n rc=$?"
  python - <<PY
import json
for what in ("bench", "rowpart", "sweep", "synth"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s_n$n.json" % what).read().strip().splitlines()[-1])
        keys = [k for k in ("value", "ms_per_step", "ms_sweep_max_rank", "load_imbalance_max_over_min", "fwd_ms", "bwd_ms", "scaling") if k in d]
        print(what, $n, {k: d[k] for k in keys})
    except Exception as e:
        print(what, $n, "failed:", e)
PY
done
