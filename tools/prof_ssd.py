"""Per-kernel CUDA time of one mamba_chunk_scan_combined call at the 9B geometry (torch profiler): python tools/prof_ssd.py [L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from tests.test_gpu_ops import _ssd_inputs
from torch.profiler import profile, ProfilerActivity
L = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, 128, 80, 8, 128, torch.bfloat16, seed=1)
run = lambda: tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        run()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"{e.key[:70]:70s} {e.device_time_total / e.count:9.1f} us x{e.count}")
