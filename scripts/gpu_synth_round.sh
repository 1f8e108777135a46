#!/bin/bash
# gpurun helper: synthesis parity tests, synthesis bench and a launch list of one batched synthesis call
mkdir -p gpurun_out
tag=${1:-a}
timeout 900 python -m pytest tests/test_gpu_synthesis.py tests/test_gpu_dio.py -x -q 2>&1 | tail -15 > gpurun_out/synth_test_$tag.log
cat gpurun_out/synth_test_$tag.log
timeout 600 python scripts/gpu_synth_bench.py 256 > gpurun_out/synth_bench_$tag.log 2>&1
tail -2 gpurun_out/synth_bench_$tag.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_synth_$tag.csv \
    python scripts/gpu_synth_bench.py 256 profile > gpurun_out/ncu_synth_$tag.log 2>&1
grep -o 'b2w::[a-z_0-9]*kernel[^"]*".*' gpurun_out/launches_synth_$tag.csv | awk -F'"' '{print $1, $(NF-1)}' | sed 's/<.*>//; s/(.*)//' | head -20
