#!/bin/bash
# developer sweep: conv parity tests + conv / norm timing for every variant library under timeviper_b200/variants
mkdir -p gpurun_out
: > gpurun_out/conv_sweep.log
for v in timeviper_b200/variants/*.so; do
  echo $v >> gpurun_out/conv_sweep.log
  TV_LIB_PATH=$PWD/$v timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k conv 2>&1 | tail -1 >> gpurun_out/conv_sweep.log
  for L in 131072 16384; do TV_LIB_PATH=$PWD/$v python tools/run_mem_kernels.py $L 10 2>&1 | head -1 >> gpurun_out/conv_sweep.log; done
done
echo base >> gpurun_out/conv_sweep.log; for L in 131072 16384; do python tools/run_mem_kernels.py $L 10 2>&1 | head -1 >> gpurun_out/conv_sweep.log; done
