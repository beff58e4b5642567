#!/bin/bash
# quick bench + full ncu capture of one kernel: gpu_r2_ncu.sh <tag> <kernel regex> [skip] [count] [extra bench args]
set -u
mkdir -p gpurun_out
TAG=$1; K=$2; SKIP=${3:-1}; CNT=${4:-1}; shift; shift; shift || true; shift || true
( XNB_TILE_DEBUG=1 timeout 600 python bench.py --steps ${STEPS:-50} --warmup 5 --no-cpu-baseline --no-e2e --no-extra "$@" ) > gpurun_out/${TAG}_bench.log 2>&1
grep "nbh_bits" gpurun_out/${TAG}_bench.log | sort | uniq -c | head -5
tail -1 gpurun_out/${TAG}_bench.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('value %.4g  ms/step %.4f  rebuilds %s' % (d['value'], d['ms_per_step'], d['config']['rebuilds'])); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()}); print('force kernel ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'whole', round(d['roofline']['whole_step']['frac'],4))
except Exception as e: print('bench failed:', l[-2000:])
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
