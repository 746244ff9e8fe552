#!/bin/bash
# time the SSD op at 128K for the default build and every variant library under timeviper_b200/variants
L=${1:-131072}
echo "default"; python tools/run_ssd.py $L 5 2>&1 | head -1
for v in timeviper_b200/variants/*.so; do echo $(basename $v); TV_LIB_PATH=$PWD/$v python tools/run_ssd.py $L 5 2>&1 | head -1; done
