#!/bin/bash
# round-2 ncu evidence (1 GPU): launch list of the bench command + one --set full capture of each path kernel
# (-> gpurun_out/; tools/ncu_traffic.py + the copies under profiles/ are made on the build box)
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ssd_fused|conv1d_fwd|gated_rmsnorm|dt_cumsum" \
    -s 8 -c 4 -f -o gpurun_out/path_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/path_${TAG}.ncu-rep gpurun_out/launches_${TAG}.csv
