#!/bin/bash
set -u
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set pdl=0 > $OUT/${TAG}_bench_nopdl.json 2> $OUT/${TAG}_bench_nopdl.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --variant ntsc_conv > $OUT/${TAG}_bench_conv.json 2> $OUT/${TAG}_bench_conv.err
python bench.py --impl dropin --dropin-seconds 3 > $OUT/${TAG}_dropin.json 2>&1
tail -2 $OUT/${TAG}_tests.log
