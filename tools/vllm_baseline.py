"""GPU baseline beside ours: vLLM's Triton port of the mamba_ssm SSD kernels and gated RMSNorm (library code
in this image, adapted from mamba_ssm v2.2.4 -- the closest available stand-in for the wheels the reference imports), on the
same inputs at the 9B geometry.  Prints parity (ours vs the port) and CUDA-event times.
    python tools/vllm_baseline.py [L] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from tests.test_gpu_ops import _ssd_inputs, relerr

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H, P, G, N, Q = 128, 80, 8, 128, 128
t_imp = time.time()
from vllm.model_executor.layers.mamba.ops.ssd_combined import mamba_chunk_scan_combined_varlen
from vllm.model_executor.layers.mamba.ops.layernorm_gated import rms_norm_gated
print(f"imported vLLM ops in {time.time() - t_imp:.1f} s", flush=True)

x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, H, P, G, N, torch.bfloat16, seed=31)
nchunks = (L + Q - 1) // Q
cu_seqlens = torch.tensor([0, L], dtype=torch.int32, device="cuda")
cu_chunk = torch.tensor(list(range(0, L, Q)) + [L], dtype=torch.int32, device="cuda")
last_chunk = torch.tensor([nchunks - 1], dtype=torch.int32, device="cuda")
seq_idx = torch.zeros(nchunks, dtype=torch.int32, device="cuda")


def timeit(fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out_v = torch.empty(L, H, P, dtype=torch.bfloat16, device="cuda")
def vllm_ssd():
    return mamba_chunk_scan_combined_varlen(x[0], dt[0], A, B[0], C[0], Q, cu_seqlens, cu_chunk, last_chunk, seq_idx, out_v,
                                            D=D.float(), z=None, dt_bias=dt_bias.float(), dt_softplus=True,
                                            state_dtype=torch.float32)
def ours_ssd():
    return tv.mamba_chunk_scan_combined(x, dt, A, B, C, Q, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)

st_v = vllm_ssd(); torch.cuda.synchronize()
y_o, st_o = ours_ssd(); torch.cuda.synchronize()
print(f"L={L}: SSD parity ours vs vLLM Triton port: y {relerr(y_o[0], out_v):.2e}  final state {relerr(st_o[0], st_v[0].float()):.2e}")
ms_v, ms_o = timeit(vllm_ssd), timeit(ours_ssd)
print(f"L={L}: SSD  vLLM Triton (5 kernels) {ms_v:.3f} ms   ours (cumsum + fused tcgen05) {ms_o:.3f} ms   speed-up {ms_v / ms_o:.2f}x")

gate = torch.randn(1, L, H * P + 64, device="cuda").to(torch.bfloat16)[..., :H * P]
w = (1 + 0.1 * torch.randn(H * P, device="cuda")).to(torch.bfloat16)
y2 = y_o.view(1, L, H * P)
n_v = rms_norm_gated(y2, w, None, z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False)
n_o = tv.rmsnorm_fn(y2, w, None, z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False)
print(f"L={L}: gated RMSNorm parity ours vs vLLM Triton: {relerr(n_o, n_v):.2e}")
ms_v = timeit(lambda: rms_norm_gated(y2, w, None, z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False))
ms_o = timeit(lambda: tv.rmsnorm_fn(y2, w, None, z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False))
print(f"L={L}: norm vLLM Triton {ms_v:.3f} ms   ours {ms_o:.3f} ms   speed-up {ms_v / ms_o:.2f}x")
