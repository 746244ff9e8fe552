"""Critical-path search by ablation (needs a --trace build): time the fused SSD kernel with parts of the work skipped."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from timeviper_b200 import _lib
from tests.test_gpu_ops import _ssd_inputs
L = 32768
x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, 128, 80, 8, 128, torch.bfloat16, seed=1)
run = lambda: tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
names = {0: "full", 2048: "full, no y stores", 1: "no epilogue (WG_C)", 2: "no M build (WG_A/H)", 4: "no decay/S-copy (WG_B)", 16: "no G MMAs", 32: "no O MMAs",
         64: "no D MMAs", 128: "no S MMAs", 240: "no MMAs at all", 7: "no WG_A/B/C work", 247: "nothing but TMA + barriers",
         247 + 256: "skeleton, no x tile TMA", 247 + 512: "skeleton, no B/C TMA", 247 + 1024: "skeleton, no L2 prefetch",
         247 + 256 + 512 + 1024: "barriers + cs/dt bulk loads only", 1024: "full, no L2 prefetch", 256: "full, no x tile TMA", 512: "full, no B/C TMA", 2048: "full, no y stores"}
for mask, nm in names.items():
    _lib.load().tv_debug_set_ablate(mask)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{nm:32s} {ms:.3f} ms  {ms * 1e3 / (L / 128):.2f} us/chunk")
_lib.load().tv_debug_set_ablate(0)
