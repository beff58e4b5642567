#!/bin/bash
# full ncu capture of one kernel of the bench step loop: gpu_ncu_kernel.sh <tag> <kernel regex> [skip] [count]
set -u
mkdir -p gpurun_out
TAG=$1; K=$2; SKIP=${3:-1}; CNT=${4:-1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
ls -la gpurun_out | tail -5
