#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest_gpu.txt 2>&1; tail -6 gpurun_out/r02e_pytest_gpu.txt
python bench.py > gpurun_out/r02e_bench_full.log 2>&1; tail -c 600 gpurun_out/r02e_bench_full.log
