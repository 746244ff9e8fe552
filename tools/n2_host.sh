#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded_host.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/n2_host_tests.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 6 > gpurun_out/n2_bench_128k.log 2>&1
tail -3 gpurun_out/n2_host_tests.log
