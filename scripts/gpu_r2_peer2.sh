#!/bin/bash
# N=2: boundary tiles on the halo stream (default) against boundary tiles on the main stream; parity first
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "2-peer" > gpurun_out/peer2_pytest.log 2>&1
echo "multi pytest rc=$?" >> gpurun_out/peer2_pytest.log
for mode in comm main comm2; do
  if [ $mode = main ]; then export XNB_BOUNDARY_ON_MAIN=1; else unset XNB_BOUNDARY_ON_MAIN; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/peer2_bench_n2_$mode.log 2>&1
  echo "bench $mode rc=$?" >> gpurun_out/peer2_bench_n2_$mode.log
done
tail -3 gpurun_out/peer2_pytest.log
