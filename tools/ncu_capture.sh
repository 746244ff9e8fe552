#!/bin/bash
# round-end ncu evidence (1 GPU): launch list of the bench command + one --set full capture of each path kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ssd_fused|conv1d_fwd|gated_rmsnorm|dt_cumsum" \
    -s 8 -c 4 -f -o gpurun_out/path_r01b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/path_r01b.ncu-rep gpurun_out/launches_r01b.csv
