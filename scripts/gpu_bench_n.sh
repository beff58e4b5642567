#!/bin/bash
# weak-scaling bench at N ranks, optionally with env settings: gpu_bench_n.sh <tag> <N> [ENV=VAL ...]
set -u
mkdir -p gpurun_out
TAG=$1; N=$2; shift 2
( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 30 --warmup 5 ) > gpurun_out/${TAG}_bench.log 2>&1
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"breakdown_ms_per_step": {[^}]*}' gpurun_out/${TAG}_bench.log | head -3
