#!/bin/bash
# fused next first half (k_lj_sweep_cl MODE 2): parity tests, then the N=1 bench with and without it
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not full_size" > gpurun_out/fused_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fused_pytest.log
for mode in fused plain fused2; do
  if [ $mode = plain ]; then export XNB_NO_FUSED_FIRST_HALF=1; else unset XNB_NO_FUSED_FIRST_HALF; fi
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/fused_bench_$mode.log 2>&1
  echo "rc=$?" >> gpurun_out/fused_bench_$mode.log
done
tail -4 gpurun_out/fused_pytest.log
