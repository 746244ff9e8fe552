#!/bin/bash
# times the conv / norm kernels for every variant library under timeviper_b200/variants (developer sweep)
mkdir -p gpurun_out
: > gpurun_out/conv_sweep.log
for v in timeviper_b200/variants/*.so; do echo $v >> gpurun_out/conv_sweep.log; for L in 131072 16384; do TV_LIB_PATH=$PWD/$v python tools/run_mem_kernels.py $L 10 >> gpurun_out/conv_sweep.log 2>&1; done; done
echo base >> gpurun_out/conv_sweep.log; python tools/run_mem_kernels.py 131072 10 >> gpurun_out/conv_sweep.log 2>&1
