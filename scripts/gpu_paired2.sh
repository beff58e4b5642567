#!/bin/bash
# bench variants of the pair-merged lists: gpu_paired2.sh <tag> "ENV=VAL ..." ["ENV=VAL ..." ...]
set -u
mkdir -p gpurun_out
TAG=$1; shift
n=0
for V in "$@"; do
n=$((n+1))
( env $V XNB_TILE_DEBUG=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e ) > gpurun_out/${TAG}_bench_$n.log 2>&1
echo "== $V"; grep "^\[xnb\] compiled" gpurun_out/${TAG}_bench_$n.log | sort | uniq -c | head -1
tail -1 gpurun_out/${TAG}_bench_$n.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('value %.4g  ms/step %.4f' % (d['value'], d['ms_per_step']), {k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()})
except Exception as e: print('bench failed:', l[-1500:])
"
done
