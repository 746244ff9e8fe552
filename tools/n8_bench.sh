#!/bin/bash
# scaling run on one 8-GPU box: bench.py at N = 8 + device timeline of one step
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench.log 2>&1
NCCL_PROTO=LL128 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/profile_sharded.py 131072 > gpurun_out/n8_profile.log 2>&1
