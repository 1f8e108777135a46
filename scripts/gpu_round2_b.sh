#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_analysis.py -m gpu -q > gpurun_out/r02b_pytest_analysis.txt 2>&1; tail -5 gpurun_out/r02b_pytest_analysis.txt
python scripts/gpu_kbench.py --utts 512 --kernels d4c,d4c_f64 > gpurun_out/r02b_kbench.txt 2>&1; cat gpurun_out/r02b_kbench.txt
ncu --set full --import-source on --clock-control none -k regex:"d4c_fast_kernel" -c 1 -o gpurun_out/prof_r02b_d4cfast python scripts/gpu_kbench.py --utts 128 --kernels d4c --reps 1 > gpurun_out/r02b_ncu.log 2>&1; tail -3 gpurun_out/r02b_ncu.log
