#!/bin/bash
# Round-2 profile bundle (run under gpurun, ONE GPU): the bench line itself (never under a profiler), the ncu launch list of the same
# command at a smaller corpus, full ncu captures of the dominant kernels, racecheck on the synthesis / D4C kernels.  Outputs in gpurun_out/.
tag=${1:-r02t}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full_$tag.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.log 2>&1
B="python bench.py --utts 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-utts 0"
$B > gpurun_out/bench_small_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"b2w|cheaptrick|d4c|mcep|lf0_vuv|bap_from|stats_kernel|render|overlap|phase|pulse|mc2sp|decode_ap|allpass" -c 600 --csv \
    --log-file gpurun_out/launches_$tag.csv $B --no-workloads > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"cheaptrick_fast_kernel|mcep_tc_kernel|d4c_fast_kernel" --launch-skip 6 -c 3 \
    -o gpurun_out/prof_${tag}_analysis $B --no-workloads > gpurun_out/ncu_full_a_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"render_fast_kernel|overlap_add_kernel|mc2sp_tc_kernel|decode_ap_kernel" --launch-skip 4 -c 4 \
    -o gpurun_out/prof_${tag}_synthesis $B --no-workloads > gpurun_out/ncu_full_s_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"allpass_tc_forward_kernel|allpass_tc_backward_kernel" --launch-skip 1 -c 3 \
    -o gpurun_out/prof_${tag}_vtln $B > gpurun_out/ncu_full_v_$tag.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_synthesis.py -m gpu -q -k "fast_render or pulse_positions or feature_domain" \
    > gpurun_out/racecheck_synthesis_$tag.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_analysis.py -m gpu -q -k "d4c_fast_path or 22k_and_48k" \
    > gpurun_out/racecheck_analysis_$tag.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_vtln.py tests/test_gpu_analysis.py -m gpu -q -k "tc or tensor or mcep or fused or mc2sp" \
    > gpurun_out/memcheck_tc_$tag.log 2>&1
tail -c 400 gpurun_out/bench_full_$tag.log; tail -3 gpurun_out/racecheck_synthesis_$tag.log gpurun_out/racecheck_analysis_$tag.log gpurun_out/memcheck_tc_$tag.log; ls -la gpurun_out/*$tag*
