"""Decode-step operators at the 9B geometry (batch 1): CUDA-event time per launch and achieved bytes/s.
python tools/run_decode.py [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
b, H, P, G, N, K = 1, 128, 80, 8, 128, 4
conv_dim = H * P + 2 * G * N
dt_ = torch.bfloat16
proj = torch.randn(b, H * P + conv_dim + H, device="cuda").to(dt_)
gate, xBC, dt = proj.split([H * P, conv_dim, H], dim=-1)
conv_state = torch.randn(b, conv_dim, K, device="cuda").to(dt_)
w = torch.randn(conv_dim, K, device="cuda").to(dt_); bias = torch.randn(conv_dim, device="cuda").to(dt_)
state = torch.randn(b, H, P, N, device="cuda")
A = (-torch.arange(1, H + 1, device="cuda").float())[:, None, None].expand(H, P, N)
D = torch.ones(H, device="cuda")[:, None].expand(H, P); dtb = torch.full((H,), -2.0, device="cuda")[:, None].expand(H, P)
nw = torch.ones(H * P, device="cuda", dtype=dt_)


def timeit(fn, nbytes, name):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name}: {us:.2f} us/launch  ({nbytes / us / 1e3:.0f} GB/s of {nbytes / 1e6:.2f} MB)")


xc = tv.causal_conv1d_update(xBC, conv_state, w, bias, "silu")
x, Bm, Cm = xc.split([H * P, G * N, G * N], dim=-1)
timeit(lambda: tv.causal_conv1d_update(xBC, conv_state, w, bias, "silu"), conv_dim * 2 * (2 + 2 * K), "causal_conv1d_update")
timeit(lambda: tv.selective_state_update(state, x.view(b, H, P), dt[:, :, None].expand(b, H, P), A, Bm.view(b, G, N),
                                         Cm.view(b, G, N), D, z=None, dt_bias=dtb, dt_softplus=True),
       2 * state.numel() * 4, "selective_state_update (fp32 state, read + write)")
y = torch.randn(b, H * P, device="cuda").to(dt_)
timeit(lambda: tv.rmsnorm_fn(y, nw, None, z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False), 3 * H * P * 2,
       "gated rmsnorm (1 row)")

# whole decode step of one layer (in_proj + conv update + state update + norm + out_proj): eager vs one CUDA graph replay
import types
cfg = tv.Mamba2Config.nanov2_9b()
mixer = tv.Mamba2MixerPrefill(cfg).to(dt_).cuda()
cache = types.SimpleNamespace(conv_states=[conv_state.clone()], ssm_states=[state.clone()], conv_kernel_size=K)
tok = torch.randn(1, 1, cfg.hidden_size, device="cuda").to(dt_)
with torch.no_grad():
    timeit(lambda: mixer.decode_step(tok, cache), 2 * state.numel() * 4, "mixer.decode_step, eager (5 kernels + glue)")
    timeit(lambda: mixer.decode_step_graph(tok, cache), 2 * state.numel() * 4, "mixer.decode_step_graph (one graph replay)")
