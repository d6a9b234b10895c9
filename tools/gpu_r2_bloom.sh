#!/bin/bash
set -u
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_bloom.py -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python bench.py --variant ntsc_bloom --steps 10 --warmup 3 --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench_ntsc_bloom.json 2> $OUT/${TAG}_bench_ntsc_bloom.err
tail -3 $OUT/${TAG}_tests.log
