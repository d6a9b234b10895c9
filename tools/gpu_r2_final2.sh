#!/bin/bash
# round 2, last capture on one GPU: whole -m gpu suite, smoke, the reference arm, the bench line of every variant, A/B lines,
# launch list, --set full captures of the hot kernels (IIR, FIR and VHS builds).  Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_gpu.csv 2>&1
python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
for v in ntsc_conv ntsc_conv4 vhs nes nes_p0 snes nesrgb template pv1k ntsc_bloom; do
    python bench.py --variant $v --steps 10 --warmup 3 --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set lines2=0 --set host_rows=0 > $OUT/${TAG}_bench_r1paths.json 2> $OUT/${TAG}_bench_r1paths.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set pdl=1 > $OUT/${TAG}_bench_pdl.json 2> $OUT/${TAG}_bench_pdl.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_launches.log 2>&1
# per step of the NTSC build: k_mod_skeleton_rgb, k_mod_picture_rgb_staged, k_sync, k_lines2 (+ empty k_lines<generic>, not matched)
ncu --set full --clock-control none --import-source on -k regex:'^k_lines2|^k_sync|^k_mod_picture_rgb_staged|^k_mod_skeleton' -s 12 -c 4 -f -o $OUT/${TAG}_ntsc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_ntsc.log 2>&1
# the FIR build launches k_lines_fir<fast> then k_lines_fir<generic> (empty) per step
ncu --set full --clock-control none --import-source on -k regex:'k_lines_fir' -s 6 -c 1 -f -o $OUT/${TAG}_ntsc_conv \
    python bench.py --variant ntsc_conv --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_noise_vhs' -s 3 -c 1 -f -o $OUT/${TAG}_vhs \
    python bench.py --variant vhs --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_vhs.log 2>&1
ls -la $OUT | grep ${TAG} | tail -40
