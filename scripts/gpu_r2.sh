#!/bin/bash
# round-2 GPU iteration: optional sanitizer pass on one small parity test, the GPU parity suite, a short bench with breakdown
set -u
mkdir -p gpurun_out
TAG=${1:-r2}; shift || true
if [ "${SANITIZE:-0}" = "1" ]; then
  ( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "${SANITIZE_K:-test_rebuild_pipeline_bit_exact and lj2k and tiled}" ) > gpurun_out/${TAG}_sanitize.log 2>&1
  echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed|at 0x" gpurun_out/${TAG}_sanitize.log | head -20
fi
( time timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -25
( XNB_TILE_DEBUG=1 timeout 600 python bench.py --steps ${STEPS:-50} --warmup 5 --no-cpu-baseline --no-e2e "$@" ) > gpurun_out/${TAG}_bench.log 2>&1
grep "^\[xnb\]" gpurun_out/${TAG}_bench.log | sort | uniq -c | head
tail -1 gpurun_out/${TAG}_bench.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('value %.4g  ms/step %.4f  rebuilds %s' % (d['value'], d['ms_per_step'], d['config']['rebuilds'])); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()}); print('force kernel ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'whole', round(d['roofline']['whole_step']['frac'],4))
except Exception as e: print('bench failed:', l[-2000:])
"
