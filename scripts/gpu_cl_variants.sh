#!/bin/bash
# tile-shape sweep of the compiled-list pair sweep (bench breakdown only)
set -u
mkdir -p gpurun_out
for T in "4,2,2" "4,4,1" "2,2,2" "4,2,1" "3,3,2" "2,2,1" "4,3,1" "3,2,2"; do
  XNB_CL_TILE=$T XNB_TILE_DEBUG=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v.log 2>&1
  grep "compiled lists" gpurun_out/v.log | head -1 | cut -c1-120
  tail -1 gpurun_out/v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip()); b=d['breakdown_ms_per_step']; print('  tile $T: ms/step %.4f force %.4f nbh %.4f' % (d['ms_per_step'], b['force'], b['nbh']))"
done
