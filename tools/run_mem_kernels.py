"""Run the two HBM-bound kernels (conv1d+SiLU, gated RMSNorm) at the 9B geometry: python tools/run_mem_kernels.py [L] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv

L = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H, P, G, N = 128, 80, 8, 128
d_inner, conv_dim = H * P, H * P + 2 * G * N
proj = torch.randn(1, L, d_inner + conv_dim + H, device="cuda").to(torch.bfloat16)
gate, xBC, dt = proj.split([d_inner, conv_dim, H], dim=-1)
w = (torch.randn(conv_dim, 4, device="cuda") * 0.5).to(torch.bfloat16)
b = torch.randn(conv_dim, device="cuda").to(torch.bfloat16)
nw = torch.ones(d_inner, device="cuda", dtype=torch.bfloat16)
y = torch.randn(1, L, d_inner, device="cuda").to(torch.bfloat16)


def timeit(fn, nbytes, name):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.3f} ms  {nbytes * L / ms / 1e6:.0f} GB/s  ({nbytes * L / ms / 1e6 / 6556.5 * 100:.1f}% of 6556.5)")


timeit(lambda: tv.causal_conv1d_fn(xBC.transpose(1, 2), w, b, activation="silu"), 49152, "conv1d+silu")
timeit(lambda: tv.rmsnorm_fn(y, nw, None, z=gate, eps=1e-5, group_size=1280, norm_before_gate=False), 61440, "gated rmsnorm")
timeit(lambda: tv.rmsnorm_fn(y, nw, None, z=None, eps=1e-5, group_size=1280, norm_before_gate=False), 40960, "rmsnorm without z (gate fused upstream)")
