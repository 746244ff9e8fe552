#!/bin/bash
# 2-GPU bench at 16K tokens/GPU (the per-GPU load of the 8-GPU 128K run) and at 64K tokens/GPU, then the 1-GPU bench.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --seqlen 32768 > gpurun_out/n2_bench_32k.log 2>&1
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_128k.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/n1_bench.log 2>&1
