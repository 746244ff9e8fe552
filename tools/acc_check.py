"""bf16 tensor-core scan vs the fp32 CUDA-core kernels in the long-memory regime (|A| / 200): python tools/acc_check.py [L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from oracle import mamba2_ref as R
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = tv.Mamba2Config.nanov2_9b()
for shift in (0.0, 5.3):
    torch.manual_seed(1234)
    p = R.nemotron_random_params(cfg.hidden_size, cfg.mamba_num_heads, cfg.mamba_head_dim, cfg.n_groups, cfg.ssm_state_size, nondegenerate=False)
    p["A_log"] = p["A_log"] - shift
    mixer = tv.Mamba2MixerPrefill(cfg).to(torch.bfloat16).cuda()
    mixer.load_state_dict({k: v.to(torch.bfloat16) for k, v in p.items()}, strict=True)
    mixer.eval()
    g = torch.Generator(device="cuda").manual_seed(4321)
    hs = torch.randn(1, L, cfg.hidden_size, device="cuda", generator=g).to(torch.bfloat16)
    with torch.no_grad():
        proj = mixer.in_proj(hs)
        core = mixer.scan_core(proj).float()
        tv.ops.force_simt_default = True
        ref32 = mixer.scan_core(proj).float()
        tv.ops.force_simt_default = False
    d = (core - ref32).abs()
    print(f"A_log shift {shift}: max|tc - fp32| / max|fp32| = {float(d.max() / ref32.abs().max()):.5f}   rms rel = {float(d.pow(2).mean().sqrt() / ref32.pow(2).mean().sqrt()):.5f}")
    for a in range(0, L, L // 4):
        sl = slice(a, a + L // 4)
        print(f"   tokens {a:6d}+: {float(d[:, sl].max() / ref32[:, sl].abs().max()):.5f}")
