#!/bin/bash
# 2-GPU check of the sharded path: NCCL parity tests, bench at 16K tokens/GPU and at 64K tokens/GPU, kernel table.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/n2_tests.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --seqlen 32768 > gpurun_out/n2_bench_32k.log 2>&1
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_128k.log 2>&1
timeout 300 $TR tools/profile_sharded.py 32768 > gpurun_out/n2_profile_32k.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/n1_bench.log 2>&1
tail -3 gpurun_out/n2_tests.log
