#!/bin/bash
# peer-memory halo: parity on 2 GPUs with both transports, then the N=2 bench with each
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/peer_pytest.log 2>&1
echo "multi pytest rc=$?" >> gpurun_out/peer_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "average_neighbors or gravitational" > gpurun_out/avg_pytest.log 2>&1
echo "avg pytest rc=$?" >> gpurun_out/avg_pytest.log
for mode in peer nccl; do
  if [ $mode = nccl ]; then export XNB_GHOST_NCCL=1; else unset XNB_GHOST_NCCL; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/peer_bench_n2_$mode.log 2>&1
  echo "bench $mode rc=$?" >> gpurun_out/peer_bench_n2_$mode.log
done
tail -3 gpurun_out/peer_pytest.log; tail -3 gpurun_out/avg_pytest.log
