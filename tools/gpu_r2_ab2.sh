#!/bin/bash
set -u
TAG=${1:-r2y}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --sustained-seconds 0 --e2e-batch 8 --set mod_bulk=0 > $OUT/${TAG}_bench_cpasync.json 2> $OUT/${TAG}_bench_cpasync.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --sustained-seconds 0 --e2e-batch 8 > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
