#!/bin/bash
# quick GPU loop for the fused SSD kernel: parity tests of the tensor-core path, then timing at 128K tokens
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ssd_tc.py -x -q 2>&1 | tail -5
timeout 120 python tools/run_ssd.py 131072 10 2>&1 | tail -4
