#!/bin/bash
# sweep ablation experiments: libs under exanbody_b200/_build/abl/lib*.so (built with -DXNB_CL_ABL=n); prints the sweep time of each
set -u
mkdir -p gpurun_out
for lib in "" $(ls exanbody_b200/_build/abl/*.so 2>/dev/null); do
  echo "== ${lib:-product}"
  env XNB_HOTPATH_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('force %.4f  step %.4f' % (d['breakdown_ms_per_step']['force'], d['ms_per_step']))
except Exception as e: print('failed', e)
"
done 2>&1 | tee gpurun_out/${1:-abl}.log
