mkdir -p gpurun_out
python scripts/bench_spmm32.py 32 20 > gpurun_out/r01d_spmm32_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmm32' --launch-skip 5 --launch-count 2 -f -o gpurun_out/r01d_spmm32_fine python scripts/bench_spmm32.py 32 1 > gpurun_out/r01d_ncu1.log 2>&1
ncu -i gpurun_out/r01d_spmm32_fine.ncu-rep --page raw --csv > gpurun_out/r01d_spmm32_fine_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_rows|k_eigval_grad_shape|k_tet_geometry|k_pack_k32' --launch-count 6 -f -o gpurun_out/r01d_asm python scripts/profile_step.py 32 1 > gpurun_out/r01d_ncu2.log 2>&1
ncu -i gpurun_out/r01d_asm.ncu-rep --page raw --csv > gpurun_out/r01d_asm_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gram_sym2|k_block_gemm|k_spmm_dual|k_eigh|k_gram<' --launch-skip 150 --launch-count 14 -f -o gpurun_out/r01d_dense python scripts/profile_step.py 32 1 > gpurun_out/r01d_ncu3.log 2>&1
ncu -i gpurun_out/r01d_dense.ncu-rep --page raw --csv > gpurun_out/r01d_dense_raw.csv
cat gpurun_out/r01d_spmm32_bench.log
