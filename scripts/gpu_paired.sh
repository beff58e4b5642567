#!/bin/bash
# A/B of the pair-merged compiled lists: parity tests and a short bench with XNB_CL_PAIRED=1, then the bench without
set -u
mkdir -p gpurun_out
TAG=${1:-pp}
( XNB_CL_PAIRED=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/${TAG}_pytest.log | tail -12
for P in 1 0; do
( XNB_CL_PAIRED=$P XNB_TILE_DEBUG=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e ) > gpurun_out/${TAG}_bench_$P.log 2>&1
grep "^\[xnb\] compiled" gpurun_out/${TAG}_bench_$P.log | sort | uniq -c | head -3
tail -1 gpurun_out/${TAG}_bench_$P.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('paired=$P value %.4g  ms/step %.4f  rebuilds %s' % (d['value'], d['ms_per_step'], d['config']['rebuilds'])); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()}); print('force kernel ms', round(d['roofline']['kernel_ms'],4))
except Exception as e: print('bench failed:', l[-1500:])
"
done
