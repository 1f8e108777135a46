#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_analysis.py -m gpu -q -x -k "mcep or fused or padded" > gpurun_out/r02l_pytest.txt 2>&1; tail -3 gpurun_out/r02l_pytest.txt
timeout 300 python scripts/gpu_kbench.py --utts 512 --kernels mcep > gpurun_out/r02l_kbench.txt 2>&1; cat gpurun_out/r02l_kbench.txt
B2W_LIB=variants/libb200world_tccopy1.so timeout 300 python scripts/gpu_kbench.py --utts 512 --kernels mcep > gpurun_out/r02l_kbench_c1.txt 2>&1; cat gpurun_out/r02l_kbench_c1.txt
