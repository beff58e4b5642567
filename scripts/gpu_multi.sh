#!/bin/bash
# multi-GPU check: parity worker + weak-scaling bench at N ranks: gpu_multi.sh <tag> <N>
set -u
mkdir -p gpurun_out
TAG=$1; N=$2
nvidia-smi -L | head -8
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-400
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 ) > gpurun_out/${TAG}_bench.log 2>&1
tail -4 gpurun_out/${TAG}_bench.log | cut -c1-1500
