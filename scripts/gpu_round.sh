#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, ncu --set full of the top kernels.
# usage (from the repo root on the GPU box): bash scripts/gpu_round.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_step.py 32 1 > gpurun_out/${TAG}_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_spmm32v|k_gram_sym|k_rr_update|k_spmm_dual|k_eigh|k_block_gemm|k_cheb32_persistent|k_assemble_rows|k_eigval_grad' \
    --launch-skip 100 --launch-count 40 -f -o gpurun_out/${TAG}_full python scripts/profile_step.py 32 1 > gpurun_out/${TAG}_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_full.ncu-rep      # gpurun_out/ travels back only below 64 MiB: keep the CSV export
# kernels outside that window: assembly (first), fine-level Gram strips / Ritz update (after the nested solve's small ones), gradient (last)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_rows_tets2|k_tet_geometry|k_eigval_grad_shape' \
    -c 3 -f -o gpurun_out/${TAG}_edge python scripts/profile_step.py 32 1 > gpurun_out/${TAG}_edge.log 2>&1; echo "ncu edge rc=$?"
ncu -i gpurun_out/${TAG}_edge.ncu-rep --page raw --csv > gpurun_out/${TAG}_edge_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_edge.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gram_strip$|k_rr_update2|k_spmm_dual_z32' \
    --launch-skip 24 --launch-count 6 -f -o gpurun_out/${TAG}_dense python scripts/profile_step.py 32 1 > gpurun_out/${TAG}_dense.log 2>&1; echo "ncu dense rc=$?"
ncu -i gpurun_out/${TAG}_dense.ncu-rep --page raw --csv > gpurun_out/${TAG}_dense_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_dense.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_synth -c 5 -f -o gpurun_out/${TAG}_synth \
    python scripts/bench_synth.py 1024 256 88200 1 > /dev/null 2>&1; echo "ncu synth rc=$?"
ncu -i gpurun_out/${TAG}_synth.ncu-rep --page raw --csv > gpurun_out/${TAG}_synth_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_synth.ncu-rep
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
timeout 300 python scripts/bench_synth.py > gpurun_out/${TAG}_synth.json 2>/dev/null
timeout 300 python scripts/bench_material.py > gpurun_out/${TAG}_material.json 2>/dev/null
timeout 120 python scripts/bench_assemble.py > gpurun_out/${TAG}_assemble.json 2>/dev/null
timeout 120 python scripts/bench_dense.py > gpurun_out/${TAG}_dense.json 2>/dev/null
timeout 120 python scripts/bench_eigh.py > gpurun_out/${TAG}_eigh.json 2>/dev/null
ls -la gpurun_out
