import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from tests.test_gpu_ops import _ssd_inputs
L = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, 128, 80, 8, 128, torch.bfloat16, seed=21)
kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
o1, f1 = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, **kw)
o1b, f1b = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, **kw)
print("deterministic:", torch.equal(o1, o1b), torch.equal(f1, f1b))
for name, kw2 in (("D scalar", kw), ("no D", dict(kw, D=None))):
    a, fa = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, **kw2)
    b, fb = tv.mamba_chunk_scan_combined(x * 2, dt, A, B, C, 128, **kw2)
    ne = (b != a * 2)
    print(name, "mismatch frac", ne.float().mean().item(), "max rel", ((b.float() - 2 * a.float()).abs().max() / a.float().abs().max()).item(),
          "state mismatch", (fb != 2 * fa).float().mean().item())
    idx = ne.nonzero()
    if len(idx):
        print(" first mismatches (b,t,h,p):", idx[:5].tolist(), "heads:", idx[:, 2].unique()[:20].tolist(), "tokens%128:", (idx[:, 1] % 128).unique()[:20].tolist())
