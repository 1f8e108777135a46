#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mgc.py tests/test_gpu_analysis.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/r02f_pytest.txt 2>&1; tail -30 gpurun_out/r02f_pytest.txt
