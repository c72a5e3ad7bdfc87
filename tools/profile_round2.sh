#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun):
#   launch list of the C4 iteration, ncu --set full of the top kernels (E-step sweeps, Gram, data-flow Cholesky, back
#   substitution) and of the K2 surface search at 200k; raw CSV pages go to gpurun_out/ for tools/ncu_summary.py.
set -x
cd "$(dirname "$0")/.."
export GINGR_CUDA_GRAPH=0
B="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"estep_colsum_kernel|estep_rowsum_kernel|gram_ws_kernel|chol_df_kernel|chol_backsolve_z_kernel" -s 10 -c 5 -f -o gpurun_out/r2_top $B > gpurun_out/r2_top.log 2>&1
ncu -i gpurun_out/r2_top.ncu-rep --page raw --csv > gpurun_out/r2_top_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_surface_warp_kernel|grid_nn_kernel|grid_line_kernel" -c 4 -f -o gpurun_out/r2_k2 python tools/k2_once.py 200000 TRIANGULAR_CLOSEST_POINT > gpurun_out/r2_k2.log 2>&1
ncu -i gpurun_out/r2_k2.ncu-rep --page raw --csv > gpurun_out/r2_k2_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
