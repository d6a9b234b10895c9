#!/bin/bash
# round 2 GPU call: A/B bench (k_lines2 vs k_lines, row-sparse e2e vs whole images), the whole -m gpu suite, launch list
# and one --set full capture of the new line kernel.   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh r2b'
set -u
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_gpu.csv 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config4-frames 0 --set lines2=0 --set host_rows=0 > $OUT/${TAG}_bench_r1paths.json 2> $OUT/${TAG}_bench_r1paths.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config4-frames 0 --sustained-seconds 0 --e2e-streams 8 --e2e-batch 256 > $OUT/${TAG}_bench_e2e8.json 2> $OUT/${TAG}_bench_e2e8.err
python -m pytest tests -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
for v in vhs nes_p0 ntsc_conv; do
    python bench.py --variant $v --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lines2|k_sync' -s 6 -c 4 -f -o $OUT/${TAG}_lines2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-batch 8 --sustained-seconds 0 --config4-frames 0 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -30
