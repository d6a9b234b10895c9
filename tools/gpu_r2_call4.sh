#!/bin/bash
# round 2: the driver's scaling run in miniature -- bench.py at N = 1, 2, 4, 8 (allgather, gather_to_root, config4) and the VHS
# variant at N = 1, 2, 4, 8 (BASELINE configs[4]).   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_r2_call4.sh r2f'
set -u
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
PORT=29600
for N in 1 8 4 2; do
  for V in ntsc vhs; do
    PORT=$((PORT+1))
    EXTRA=""
    if [ "$V" = "vhs" ]; then EXTRA="--no-cpu-baseline"; fi
    if [ "$N" = "1" ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --variant $V $EXTRA > $OUT/${TAG}_bench_${V}_n1.json 2> $OUT/${TAG}_bench_${V}_n1.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --variant $V > $OUT/${TAG}_bench_${V}_n$N.json 2> $OUT/${TAG}_bench_${V}_n$N.err
    fi
    echo "N=$N $V rc=$?" >> $OUT/${TAG}_status.txt
  done
done
cat $OUT/${TAG}_status.txt
