#!/bin/bash
# round-2 checkpoint on one GPU: parity suite, default bench (with parity + extra workloads), launch list, full captures of the two large kernels
set -u
mkdir -p gpurun_out
TAG=${1:-r2z}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.log 2>&1
tail -2 gpurun_out/${TAG}_bench.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lj_sweep -s 4 -c 1 -f -o gpurun_out/${TAG}_force \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${TAG}_ncu_force.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nbh_bits -s 1 -c 1 -f -o gpurun_out/${TAG}_nbh \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${TAG}_ncu_nbh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nbh_big -s 1 -c 1 -f -o gpurun_out/${TAG}_nbh_big \
   python bench.py --workload C4 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${TAG}_ncu_nbh_big.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -1 gpurun_out/${TAG}_bench_ref.log | cut -c1-400
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
ls gpurun_out | grep ${TAG}
