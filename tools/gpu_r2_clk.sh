#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
./ntsc-crt_b200/bin/ubench_pipes > $OUT/r2n_ubench_pipes.txt 2>&1
python tools/phase_clocks.py > $OUT/r2n_phase_clocks.txt 2>&1
python tools/phase_clocks.py pdl=0 > $OUT/r2n_phase_clocks_nopdl.txt 2>&1
cat $OUT/r2n_ubench_pipes.txt
cat $OUT/r2n_phase_clocks_nopdl.txt
