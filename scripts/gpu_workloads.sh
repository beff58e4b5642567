#!/bin/bash
# bench lines of the other SURVEY 8(d) workloads (parity-test cases, not the headline): gpu_workloads.sh <tag>
set -u
mkdir -p gpurun_out
TAG=$1
for W in C1 C4 C5; do
  ( XNB_TILE_DEBUG=1 timeout 600 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/${TAG}_${W}.log 2>&1
  grep "compiled lists\|nbh tiles" gpurun_out/${TAG}_${W}.log | tail -2 | cut -c1-200
  tail -1 gpurun_out/${TAG}_${W}.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('$W value %.4g  ms/step %.4f  rebuilds %s atoms %d' % (d['value'], d['ms_per_step'], d['config']['rebuilds'], d['config']['atoms'])); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()})
except Exception as e: print('$W failed:', l[-600:])
"
done
