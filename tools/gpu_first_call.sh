#!/bin/bash
# One gpurun call that brings back everything the next round needs about the variants that have not run on a GPU yet
# (template, PV-1000, bloom, NES / NES-RGB chroma patterns, wire-format kernels) next to the flagship numbers:
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh r2a'
# Results land in gpurun_out/<tag>_* ; copy what is worth keeping to profiles/.
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_gpu.csv 2>&1

# 1. parity: the whole -m gpu suite (the never-run variants are collected last, tests/conftest.py)
python -m pytest tests -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
# ... and, whatever happened above, the late group on its own without -x, so that one failure does not hide the rest
python -m pytest tests -m gpu -q -k "template or pv1k or bloom or nes_p1 or nesrgb_p or wire or still or edges or fullsize" > $OUT/${TAG}_tests_new.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests_new.log

# 2. bench lines: flagship first (with cpu baseline), then the other variants (informational, no cpu baseline)
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
for v in ntsc_conv template pv1k ntsc_bloom nes_p1 nesrgb_p0 snes vhs; do
    python bench.py --variant $v --steps 6 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
done

# 3. launch lists (shares of the step, cold-cache and serialised -- never a bench value)
for v in ntsc pv1k ntsc_bloom template; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_$v.csv \
        python bench.py --variant $v --steps 2 --warmup 1 --no-cpu-baseline --e2e-batch 8 > $OUT/${TAG}_ncu_$v.log 2>&1
done

# 4. one --set full capture of the line / sync / encoder kernels of the PV-1000 and of the bloom decoder
ncu --set full --clock-control none --import-source on -k regex:'k_lines|k_sync|k_mod_picture_rgb' -c 6 -f -o $OUT/${TAG}_pv1k \
    python bench.py --variant pv1k --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 8 > $OUT/${TAG}_ncu_full_pv1k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lines_bloom|k_bloom' -c 4 -f -o $OUT/${TAG}_bloom \
    python bench.py --variant ntsc_bloom --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 8 > $OUT/${TAG}_ncu_full_bloom.log 2>&1
ls -la $OUT | tail -40
