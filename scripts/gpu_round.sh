#!/bin/bash
# One GPU-box call: parity tests, bench, ncu launch list and a full ncu capture of the pair sweep.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py --steps 100 --warmup 10 ) > gpurun_out/${TAG}_bench.log 2>&1
tail -2 gpurun_out/${TAG}_bench.log | cut -c1-1800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lj_sweep -s 4 -c 2 -f -o gpurun_out/${TAG}_force \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_force.log 2>&1
ls -la gpurun_out | tail -20
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -1 gpurun_out/${TAG}_bench_ref.log | cut -c1-600
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
