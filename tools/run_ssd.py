"""Run the SSD op alone at the 9B geometry (profiling helper): python tools/run_ssd.py [L] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from tests.test_gpu_ops import _ssd_inputs

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
G = int(sys.argv[4]) if len(sys.argv) > 4 else 8
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, H, 80, G, 128, torch.bfloat16, seed=1)
for _ in range(2):
    tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
zms = None
if os.environ.get("TV_RUN_Z"):
    for _ in range(2):
        tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, z=z, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, z=z, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
    e1.record(); torch.cuda.synchronize()
    print(f"with z gating fused in the epilogue: {e0.elapsed_time(e1) / iters:.3f} ms")
print(f"L={L} H={H} G={G} ssd {ms:.3f} ms  {(45312/128*H)*L/ms/1e6:.1f} GB/s  {ms*1e3/ (L/128):.2f} us/chunk")
# shard summary (pass 1 of the sequence-sharded path)
for scale, tag in ((1.0, "fast decay (A = -1..-H)"), (1e-4, "slow decay (A * 1e-4: every chunk contributes)")):
    A2 = A * scale
    for _ in range(2):
        tv.mamba_chunk_state_summary(x, dt, A2, B, 128, dt_bias=dt_bias, dt_softplus=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        tv.mamba_chunk_state_summary(x, dt, A2, B, 128, dt_bias=dt_bias, dt_softplus=True)
    e1.record(); torch.cuda.synchronize()
    print(f"  state summary, {tag}: {e0.elapsed_time(e1) / iters:.3f} ms")
