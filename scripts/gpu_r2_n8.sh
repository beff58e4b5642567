#!/bin/bash
# 8 GPUs: parity (peer transport), the C3 weak-scaling bench at N=8 and N=4 with the peer-memory halo, and the N=8 A/B against
# the NCCL transport with the boundary tiles on the main stream (the previous configuration)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "8-peer or 4-peer or 2-nccl" > gpurun_out/n8_pytest.log 2>&1
echo "multi pytest rc=$?" >> gpurun_out/n8_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/n8_bench_peer.log 2>&1
echo "rc=$?" >> gpurun_out/n8_bench_peer.log
XNB_GHOST_NCCL=1 XNB_BOUNDARY_ON_MAIN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 bench.py --gpus 8 --steps 30 --warmup 5 --no-extra > gpurun_out/n8_bench_nccl.log 2>&1
echo "rc=$?" >> gpurun_out/n8_bench_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/n4_bench_peer.log 2>&1
echo "rc=$?" >> gpurun_out/n4_bench_peer.log
tail -3 gpurun_out/n8_pytest.log
