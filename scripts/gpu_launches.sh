#!/bin/bash
# ncu launch list only: gpu_launches.sh <tag> [env...]
set -u
mkdir -p gpurun_out
TAG=$1; shift
env "$@" XNB_TILE_DEBUG=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launches.log 2>&1
grep "nbh tiles" gpurun_out/${TAG}_ncu_launches.log | head -20
