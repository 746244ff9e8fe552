#!/bin/bash
# developer sweep: parity tests + SSD timing at 128K for every variant library under timeviper_b200/variants
mkdir -p gpurun_out; : > gpurun_out/ssd_sweep.log
timeout 300 python -m pytest tests/test_gpu_ssd_tc.py -x -q -m gpu >> gpurun_out/ssd_sweep.log 2>&1
echo "default build" >> gpurun_out/ssd_sweep.log; python tools/run_ssd.py 131072 5 2>&1 | head -1 >> gpurun_out/ssd_sweep.log
for v in timeviper_b200/variants/*.so; do echo $v >> gpurun_out/ssd_sweep.log; TV_LIB_PATH=$PWD/$v python tools/run_ssd.py 131072 5 2>&1 | head -1 >> gpurun_out/ssd_sweep.log; done
