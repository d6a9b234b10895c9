#!/bin/bash
# the --set full captures of the final build (one launch of each hot kernel at the full batch of 296 monitors)
set -u
TAG=${1:-r2_final}
OUT=gpurun_out
mkdir -p $OUT
# per step of the NTSC build: k_mod_skeleton_rgb, k_mod_picture_rgb_staged, k_mod_picture_rgb (gather, empty), k_sync, k_lines2, k_lines<generic> (empty)
ncu --set full --clock-control none --import-source on -k regex:'^k_lines2|^k_sync|^k_mod_picture_rgb_staged|^k_mod_skeleton' -s 12 -c 4 -f -o $OUT/${TAG}_ntsc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_ntsc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lines_fir<1' -s 3 -c 1 -f -o $OUT/${TAG}_ntsc_conv \
    python bench.py --variant ntsc_conv --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_conv.log 2>&1
ls -la $OUT/*.ncu-rep
