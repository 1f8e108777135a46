#!/bin/bash
# ncu --set full captures of the HBM-bound / new kernels (one launch each): overlap-add, decode_ap, mc2sp (synthesis call),
# pad / un-pad (trainer batch), DIO low cut / band / StoneMask.  Outputs in gpurun_out/.
tag=${1:-r01l}
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"overlap_add_kernel|decode_ap_kernel|mc2sp_kernel" -c 3 \
    -o gpurun_out/prof_synth_hbm_$tag python scripts/gpu_synth_bench.py 256 profile > gpurun_out/ncu_synth_hbm_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pad_normalise4_kernel|unpad_denormalise4_kernel" --launch-skip 6 -c 2 \
    -o gpurun_out/prof_pad_$tag python scripts/bench_extra.py batch_only > gpurun_out/ncu_pad_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"dio_lowcut_kernel|dio_band_kernel|stonemask_kernel" --launch-skip 3 -c 3 \
    -o gpurun_out/prof_f0_$tag python scripts/gpu_f0_bench.py 128 > gpurun_out/ncu_f0_$tag.log 2>&1
ls -la gpurun_out/prof_*_$tag.ncu-rep
