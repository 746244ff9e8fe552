"""Smallest fused-SSD run with error reporting (debug helper): python tools/dbg_ssd.py [L] [H] [G]"""
import sys, os
os.environ.setdefault("CUDA_LAUNCH_BLOCKING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from tests.test_gpu_ops import _ssd_inputs, _cpu, relerr
from oracle import mamba2_ref as R
L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H = int(sys.argv[2]) if len(sys.argv) > 2 else 4
G = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, H, 80, G, 128, torch.bfloat16, seed=11)
out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
torch.cuda.synchronize()
ref, ref_fin = R.ssd_chunked_ref(*_cpu(x, dt, A, B, C), 128, D=D.cpu(), dt_bias=dt_bias.cpu(), dt_softplus=True)
print("L", L, "H", H, "out relerr", relerr(out, ref), "fin relerr", relerr(fin, ref_fin))
nch = (L + 127) // 128
o = out.float().cpu().reshape(1, L, H, 80); rf = ref.float().reshape(1, L, H, 80)
for c in range(min(nch, 6)):
    sl = slice(c * 128, min(L, (c + 1) * 128))
    print(" chunk", c, "relerr", ((o[:, sl] - rf[:, sl]).abs().max() / rf[:, sl].abs().max()).item(),
          "by row quarter", [round(((o[:, c*128+32*q:c*128+32*q+32] - rf[:, c*128+32*q:c*128+32*q+32]).abs().max()).item(), 4) for q in range(4) if c*128+32*q < L])
