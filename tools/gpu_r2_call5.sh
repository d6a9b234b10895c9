#!/bin/bash
set -u
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -x -k "vhs or video or api" > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python bench.py --variant vhs --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_vhs.json 2> $OUT/${TAG}_bench_vhs.err
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
ncu --set full --clock-control none --import-source on -k regex:'k_noise_vhs' -s 2 -c 1 -f -o $OUT/${TAG}_vhs \
    python bench.py --variant vhs --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu.log 2>&1
ls -la $OUT | tail -6
