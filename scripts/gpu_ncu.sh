#!/bin/bash
# ncu launch list + full capture of one kernel: gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
set -u
mkdir -p gpurun_out
TAG=$1; KRE=$2; SKIP=${3:-1}; CNT=${4:-1}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP} -c ${CNT} -f -o gpurun_out/${TAG}_k \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_k.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_k.log | cut -c1-300
