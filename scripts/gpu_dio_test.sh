#!/bin/bash
# gpurun helper: F0-stage parity tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dio.py -x -q 2>&1 | tail -40 > gpurun_out/dio_test.log
cat gpurun_out/dio_test.log
timeout 600 python scripts/gpu_f0_bench.py 512 > gpurun_out/f0_bench.log 2>&1
tail -3 gpurun_out/f0_bench.log
