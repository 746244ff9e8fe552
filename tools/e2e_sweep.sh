#!/bin/bash
# e2e (host-buffer streamed prefill) vs segment size
for seg in 4096 8192 32768; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-segment $seg --e2e-steps 5 > gpurun_out/b_seg_$seg.log 2>&1
  tail -1 gpurun_out/b_seg_$seg.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($seg, d['e2e']['ms_per_step'], d['e2e']['value'])"
done
