#!/bin/bash
# run the short bench under several env-variable settings: gpu_variants.sh <tag> "<VAR=val ...>" "<VAR=val ...>" ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
if [ -z "${SKIP_TESTS:-}" ]; then ( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log; fi
exec > >(tee gpurun_out/${TAG}_variants.log) 2>&1
for v in "$@"; do
  echo "== $v"
  env $v XNB_TILE_DEBUG=1 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
seen=set()
for l in sys.stdin.read().strip().splitlines():
    if l.startswith('[xnb]'):
        k=l[:14]
        if k not in seen: print(l); seen.add(k)
    if l.startswith('{'):
        d=json.loads(l); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()}, 'ms/step', round(d['ms_per_step'],4), 'rebuilds', d['config']['rebuilds'])
"
done
