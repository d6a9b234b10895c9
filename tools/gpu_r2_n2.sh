#!/bin/bash
# the final build on two GPUs, launched as the driver launches it: the bench line (allgather, gather_to_root, config4 inside), the
# VHS line, and the partition checks over NCCL.   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2_n2.sh r2u'
set -u
TAG=${1:-r2u}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err
echo "ntsc rc=$?" > $OUT/${TAG}_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 --variant vhs --no-cpu-baseline --config4-frames 0 > $OUT/${TAG}_bench_2gpu_vhs.json 2> $OUT/${TAG}_bench_2gpu_vhs.err
echo "vhs rc=$?" >> $OUT/${TAG}_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tests/multi_gpu_check.py > $OUT/${TAG}_multi_gpu_check.json 2> $OUT/${TAG}_multi_gpu_check.err
echo "check rc=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt
