"""Dump the per-chunk pipeline timeline (clock64) of CTA (0,0) of the fused SSD kernel."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from timeviper_b200 import _lib
from tests.test_gpu_ops import _ssd_inputs

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ABL = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = L // 128
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, 128, 80, 8, 128, torch.bfloat16, seed=1)
run = lambda: tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
_lib.load().tv_debug_set_ablate(ABL)
run(); torch.cuda.synchronize()
buf = torch.zeros(2 * n * 16, dtype=torch.int64, device="cuda")
_lib.load().tv_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
_lib.load().tv_debug_set_trace(None)
tall = buf.cpu().view(2 * n, 16)
t = tall[:n]
ta = tall[n:]
names = ["P:x issue", "I:G issue", "I:S issue", "I:O issue", "I:D issue", "A:cbfull", "A:m_done", "X:x landed",
         "X:stdone(c-1)", "X3:xs written", "I:passes", "C:yfull(c)", "C:acc freed", "C:epi_done(c)", "S:fold+copy"]
t0 = int(t[t > 0].min())
c0, c1 = n // 2, n // 2 + 4
print("chunk | " + " | ".join(f"{nm:>14}" for nm in names))
for c in range(c0, c1):
    print(f"{c:5d} | " + " | ".join(f"{int(t[c, e]) - t0:14d}" for e in range(len(names))))
dcy = (t[n - 1, 0] - t[2, 0]).item(); dns = (t[n - 1, 15] - t[2, 15]).item()
print(f"SM clock during the kernel: {dcy / dns * 1e3:.0f} MHz over {dns/1e3:.1f} us; {dcy/(n-3):.0f} cycles/chunk average")
d = (t[c1, 2] - t[c0, 2]).item() / (c1 - c0)
print("cycles per chunk (S issue to S issue):", d)
for e, nm in enumerate(names):
    rel = (t[c0:c1, e] - t[c0:c1, 2]).float().mean().item()
    print(f"  {nm:>16} relative to S issue of same chunk: {rel:10.0f}")

