#!/bin/bash
# phase clocks of the debug build + one bench line + the parity tests closest to the kernels (a quick look between changes)
set -u
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
python tools/phase_clocks.py pdl=0 > $OUT/${TAG}_phase_clocks.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set pdl=0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_lines2.py -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
cat $OUT/${TAG}_phase_clocks.txt
tail -3 $OUT/${TAG}_tests.log
