#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_analysis.py -m gpu -q > gpurun_out/r02j_pytest.txt 2>&1; tail -6 gpurun_out/r02j_pytest.txt
python scripts/gpu_kbench.py --utts 512 --kernels cheaptrick > gpurun_out/r02j_kbench.txt 2>&1; cat gpurun_out/r02j_kbench.txt
ncu --set full --import-source on --clock-control none -k regex:"cheaptrick_fast_kernel" -c 1 -o gpurun_out/prof_r02j_ctfast python scripts/gpu_kbench.py --utts 128 --kernels cheaptrick --reps 1 > gpurun_out/r02j_ncu.log 2>&1; tail -2 gpurun_out/r02j_ncu.log
