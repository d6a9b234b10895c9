#!/bin/bash
set -u
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
for v in ntsc nes nesrgb snes pv1k; do
    python bench.py --variant $v --steps 10 --warmup 3 --no-cpu-baseline --config4-frames 0 --sustained-seconds 0 --e2e-batch 8 > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
tail -3 $OUT/${TAG}_tests.log
