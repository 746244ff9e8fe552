#!/bin/bash
# one 8-GPU box: sharded parity tests at the timed configuration (tracked log), bench at N = 8, device timeline of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_sharded_host.py -q -m gpu -s > gpurun_out/r02_n8_parity_tests.log 2>&1
tail -4 gpurun_out/r02_n8_parity_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_n8_bench.log 2>&1
tail -1 gpurun_out/r02_n8_bench.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/profile_sharded.py 131072 > gpurun_out/r02_n8_profile.log 2>&1
tail -3 gpurun_out/r02_n8_profile.log
