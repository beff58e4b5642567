#!/bin/bash
# bench-only comparison of environment variants: gpu_r2_variants.sh <tag> "VAR=1 VAR2=x" "VAR=2" ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
for v in "$@"; do
  echo "== $v"
  ( env $v XNB_TILE_DEBUG=1 timeout 300 python bench.py --steps ${STEPS:-50} --warmup 5 --no-cpu-baseline --no-e2e --no-extra ) > gpurun_out/${TAG}_v.log 2>&1
  grep "nbh_bits" gpurun_out/${TAG}_v.log | tail -1 | cut -c1-200
  tail -1 gpurun_out/${TAG}_v.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('value %.4g  ms/step %.4f  rebuilds %s' % (d['value'], d['ms_per_step'], d['config']['rebuilds']), {k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()})
except Exception as e: print('bench failed:', l[-1500:])
"
done
