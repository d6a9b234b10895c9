#!/bin/bash
# A/B timings of one build: default, encoder staging by cp.async, tests of the touched kernels
set -u
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_lines2.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set mod_bulk=0 > $OUT/${TAG}_bench_modcp.json 2> $OUT/${TAG}_bench_modcp.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --variant vhs > $OUT/${TAG}_bench_vhs.json 2> $OUT/${TAG}_bench_vhs.err
tail -2 $OUT/${TAG}_tests.log
