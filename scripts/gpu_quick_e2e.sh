#!/bin/bash
# quick GPU iteration: parity tests + a short bench WITH the host-buffer leg (no CPU baseline); extra args go to bench.py
set -u
mkdir -p gpurun_out
TAG=${1:-q}; shift || true
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -12
( timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline "$@" ) > gpurun_out/${TAG}_bench.log 2>&1
tail -1 gpurun_out/${TAG}_bench.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('value %.4g  ms/step %.4f  rebuilds %s' % (d['value'], d['ms_per_step'], d['config']['rebuilds'])); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()}); print('force kernel ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4)); print('e2e', d['e2e'])
except Exception as e: print('bench failed:', l[-2000:])
"
