#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r02c_pytest_gpu.txt
python bench.py --utts 600 --steps 2 --warmup 1 > gpurun_out/r02c_bench_small.log 2>&1; tail -c 3000 gpurun_out/r02c_bench_small.log
