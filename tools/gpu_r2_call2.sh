#!/bin/bash
# round 2: k_lines2 v2.1 A/B (staging by TMA vs cp.async), tests of the kernel, ncu capture.
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_lines2.py tests/test_gpu_parity.py tests/test_gpu_batch_api.py tests/test_gpu_fullsize.py tests/test_gpu_video_convert_unmodified.py tests/test_gpu_lineshard.py -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench_tma1.json 2> $OUT/${TAG}_bench_tma1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set tma=2 > $OUT/${TAG}_bench_tma2.json 2> $OUT/${TAG}_bench_tma2.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set tma=2 --e2e-streams 8 --e2e-batch 256 --sustained-seconds 0 > $OUT/${TAG}_bench_e2e8.json 2> $OUT/${TAG}_bench_e2e8.err
ncu --set full --clock-control none --import-source on -k regex:'k_lines2' -s 3 -c 1 -f -o $OUT/${TAG}_lines2_tma1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lines2' -s 3 -c 1 -f -o $OUT/${TAG}_lines2_tma2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 --set tma=2 > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT | tail -12
