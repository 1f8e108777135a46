#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_synthesis.py tests/test_gpu_pipeline.py -m gpu -q -k "not baseline_sized" > gpurun_out/r02v_memcheck_synth.log 2>&1; tail -4 gpurun_out/r02v_memcheck_synth.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_synthesis.py -m gpu -q -k "pulse_positions or feature_domain or fast_render" > gpurun_out/r02v_racecheck_synth.log 2>&1; tail -3 gpurun_out/r02v_racecheck_synth.log
