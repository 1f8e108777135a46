#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vtln.py -m gpu -q -x > gpurun_out/r02n_pytest.txt 2>&1; tail -3 gpurun_out/r02n_pytest.txt
timeout 300 python scripts/gpu_vtln_bench.py bwd > gpurun_out/r02n_vtln.txt 2>&1; cat gpurun_out/r02n_vtln.txt
