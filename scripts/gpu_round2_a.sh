#!/bin/bash
# round 2, first GPU call: parity of the single-precision D4C path, kernel timings of variants, one full ncu capture
mkdir -p gpurun_out
python -m pytest tests/test_gpu_analysis.py -m gpu -x -q > gpurun_out/r02a_pytest_analysis.txt 2>&1; tail -5 gpurun_out/r02a_pytest_analysis.txt
python scripts/gpu_kbench.py --utts 512 --kernels cheaptrick,mcep,d4c,d4c_f64 > gpurun_out/r02a_kbench.txt 2>&1; cat gpurun_out/r02a_kbench.txt
for v in d4cf4 d4cf6; do B2W_LIB=variants/libb200world_$v.so python scripts/gpu_kbench.py --utts 512 --kernels d4c > gpurun_out/r02a_kbench_$v.txt 2>&1; tail -4 gpurun_out/r02a_kbench_$v.txt; done
ncu --set full --import-source on --clock-control none -k regex:"d4c_fast_kernel" -c 1 -o gpurun_out/prof_r02a_d4cfast python scripts/gpu_kbench.py --utts 128 --kernels d4c --reps 1 > gpurun_out/r02a_ncu.log 2>&1; tail -3 gpurun_out/r02a_ncu.log
