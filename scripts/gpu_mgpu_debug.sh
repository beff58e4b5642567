#!/bin/bash
# run the multi-GPU parity worker at N ranks with full output kept: gpu_mgpu_debug.sh <tag> <N>
set -u
mkdir -p gpurun_out
TAG=$1; N=$2
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tests/mgpu_worker.py ) > gpurun_out/${TAG}_worker.log 2>&1
grep -n "parity ok\|Error\|error\|assert\|differs\|lost" gpurun_out/${TAG}_worker.log | head -30 | cut -c1-400
