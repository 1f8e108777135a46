#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"mlpg_solve_kernel|mlpg_factor_kernel|deltas_rows_kernel|world_metrics_rows_kernel" --launch-skip 4 -c 4 \
    -o gpurun_out/prof_r03c_post env PYTHONPATH=. python scripts/gpu_post_prof.py > gpurun_out/ncu_r03c_post.log 2>&1
tail -3 gpurun_out/ncu_r03c_post.log; ls -la gpurun_out/prof_r03c_post.ncu-rep
