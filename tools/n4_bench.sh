#!/bin/bash
# 4-GPU bench (sequence-sharded, 131072 tokens) + device timeline of one step
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.log 2>&1
NCCL_PROTO=LL128 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 tools/profile_sharded.py 131072 > gpurun_out/n4_profile.log 2>&1
