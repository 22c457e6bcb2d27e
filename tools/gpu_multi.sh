#!/bin/bash
# N-GPU bench through torchrun exactly as the driver launches it.  usage: gpu_multi.sh N [extra bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?"; tail -c 3000 gpurun_out/bench_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "ref rc=$?"; tail -c 800 gpurun_out/bench_ref_n$N.log
