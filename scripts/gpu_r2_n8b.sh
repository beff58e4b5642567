#!/bin/bash
# 8-GPU box: C3 weak scaling at N=8 and N=4 with the final code (peer-memory halo, fused first half)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/n8b_bench.log 2>&1
echo "rc=$?" >> gpurun_out/n8b_bench.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/n4b_bench.log 2>&1
echo "rc=$?" >> gpurun_out/n4b_bench.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/n2b_bench.log 2>&1
echo "rc=$?" >> gpurun_out/n2b_bench.log
