#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"mcep_tc_kernel" -c 1 -o gpurun_out/prof_r02m_mcep python scripts/gpu_kbench.py --utts 128 --kernels mcep --reps 1 > gpurun_out/r02m_ncu.log 2>&1; tail -2 gpurun_out/r02m_ncu.log
