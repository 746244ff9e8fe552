#!/bin/bash
# scaling run on one 8-GPU box: bench.py at N = 8, 4 (N = 1, 2 come from the cheaper 2-GPU box runs)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/n${n}_bench.log 2>&1
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/profile_sharded.py 131072 > gpurun_out/n8_profile.log 2>&1
