#!/bin/bash
# round-end check on one GPU: full GPU test tier, smoke(), default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/final_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.log 2>&1
tail -3 gpurun_out/final_tests.log; tail -1 gpurun_out/final_smoke.log
