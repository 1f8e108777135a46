#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_vtln.py tests/test_gpu_analysis.py -m gpu -q -k "tensor_core or mcep_vs_oracle" > gpurun_out/r02r_synccheck_tc.log 2>&1; tail -4 gpurun_out/r02r_synccheck_tc.log
timeout 300 python scripts/gpu_kbench.py --utts 512 --kernels mcep > gpurun_out/r02l_kbench.txt 2>&1; cat gpurun_out/r02l_kbench.txt
