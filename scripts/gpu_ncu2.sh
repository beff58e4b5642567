#!/bin/bash
# full ncu captures of several kernels in one bench run: gpu_ncu2.sh <tag> <regex1> <regex2> ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_${K} \
     python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_${K}_ncu.log 2>&1
  tail -1 gpurun_out/${TAG}_${K}_ncu.log | cut -c1-200
done
( XNB_TILE_DEBUG=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e ) 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip()); print('value %.4g  ms/step %.4f  rebuilds %s' % (d['value'], d['ms_per_step'], d['config']['rebuilds'])); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()})"
ls -la gpurun_out | tail -8
