#!/bin/bash
# N-GPU sweep of NCCL settings for the boundary-state all-gather:  tools/ag_sweep.sh N
N=${1:-8}
mkdir -p gpurun_out; : > gpurun_out/ag_sweep.log
run() { env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/ag_bench.py 2>&1 | grep "all-gather of" >> gpurun_out/ag_sweep.log; }
run NCCL_DEBUG=WARN
run NCCL_ALGO=NVLS
run NCCL_PROTO=Simple
run NCCL_PROTO=LL128
run NCCL_MIN_NCHANNELS=32
run NCCL_ALGO=Ring NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32
cat gpurun_out/ag_sweep.log
