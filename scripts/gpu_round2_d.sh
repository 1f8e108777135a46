#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_synthesis.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/r02d_pytest_synth.txt 2>&1; tail -15 gpurun_out/r02d_pytest_synth.txt
python bench.py --utts 600 --steps 2 --warmup 1 --no-workloads > gpurun_out/r02d_bench_small.log 2>&1; tail -c 1500 gpurun_out/r02d_bench_small.log
