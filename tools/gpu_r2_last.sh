#!/bin/bash
# the last call of round 2: whole -m gpu suite, the PV-1000 line (staged encoder), the contract line and the --set full capture
# of the NTSC kernels of the same build (profiles/make_traffic.py ties the DRAM traffic to the source hash)
set -u
TAG=${1:-r2last}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python bench.py --variant pv1k --steps 10 --warmup 3 --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench_pv1k.json 2> $OUT/${TAG}_bench_pv1k.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
ncu --set full --clock-control none --import-source on -k regex:'^k_lines2|^k_sync|^k_mod_picture_rgb_staged|^k_mod_skeleton' -s 12 -c 4 -f -o $OUT/${TAG}_ntsc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_ntsc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lines_fir' -s 6 -c 1 -f -o $OUT/${TAG}_ntsc_conv \
    python bench.py --variant ntsc_conv --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_conv.log 2>&1
tail -3 $OUT/${TAG}_tests.log
