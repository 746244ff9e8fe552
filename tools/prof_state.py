import sys, os
sys.path.insert(0, "/root/repo")
import torch
import timeviper_b200 as tv
from tests.test_gpu_ops import _ssd_inputs
from torch.profiler import profile, ProfilerActivity
L = int(sys.argv[1])
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, 128, 80, 8, 128, torch.bfloat16, seed=1)
for _ in range(3):
    tv.mamba_chunk_state_summary(x, dt, A, B, 128, dt_bias=dt_bias, dt_softplus=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        tv.mamba_chunk_state_summary(x, dt, A, B, 128, dt_bias=dt_bias, dt_softplus=True)
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(f"{e.key[:60]:60s} {e.device_time_total / e.count:8.1f} us x{e.count}")
