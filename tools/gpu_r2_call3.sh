#!/bin/bash
# round 2: two GPUs -- the N > 1 paths of bench.py exactly as the driver launches them (allgather, gather_to_root, config4 over
# frame ranges), the multi-GPU correctness script, VHS at N = 2; then on GPU 0: PCIe micro-benchmark, anchors, whole suite.
set -u
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --variant vhs > $OUT/${TAG}_bench_2gpu_vhs.json 2> $OUT/${TAG}_bench_2gpu_vhs.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_check.py > $OUT/${TAG}_multi_gpu_check.log 2>&1
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
./ntsc-crt_b200/bin/ubench_pcie > $OUT/${TAG}_ubench_pcie.txt 2>&1
python -m pytest tests -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
ls -la $OUT | tail -12
