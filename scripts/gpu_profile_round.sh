#!/bin/bash
# Round profile bundle (run under gpurun): launch list of the bench command, full ncu capture of the three main kernels,
# the bench line itself (not under a profiler) and the secondary measurements.  Outputs in gpurun_out/.
tag=${1:-r01h}
python bench.py > gpurun_out/bench_full_$tag.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.log 2>&1
python scripts/bench_extra.py > gpurun_out/bench_extra_$tag.log 2>&1
python scripts/gpu_vtln_bench.py bwd >> gpurun_out/bench_extra_$tag.log 2>&1
python scripts/gpu_f0_bench.py 512 >> gpurun_out/bench_extra_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cheaptrick|d4c|mcep|lf0_vuv|bap_from|stats_kernel|dio_|stonemask" -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"cheaptrick_kernel|mcep_tc_kernel|d4c_kernel" --launch-skip 3 -c 3 \
    -o gpurun_out/prof_$tag python bench.py --utts 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -c 600 gpurun_out/bench_full_$tag.log
