"""Burst vs sustained: times blocks of 20 hot-path steps back to back for ~4 s and samples SM clock / power.
python tools/sustained.py [blocks]"""
import os, sys, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv
from oracle import mamba2_ref as R

cfg = tv.Mamba2Config.nanov2_9b()
L = 131072
p = R.nemotron_random_params(cfg.hidden_size, cfg.mamba_num_heads, cfg.mamba_head_dim, cfg.n_groups, cfg.ssm_state_size, nondegenerate=False)
mixer = tv.Mamba2MixerPrefill(cfg).to(torch.bfloat16).cuda()
mixer.load_state_dict({k: v.to(torch.bfloat16) for k, v in p.items()})
proj = (torch.randn(1, L, cfg.projection_size, device="cuda") * 0.5).to(torch.bfloat16)
rows = []
pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                       "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in pr.stdout], daemon=True).start()
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 30
with torch.no_grad():
    for _ in range(3):
        mixer.scan_core(proj)
    torch.cuda.synchronize()
    time.sleep(1.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(nblocks + 1)]
    t0 = time.time()
    ev[0].record()
    for i in range(nblocks):
        for _ in range(20):
            mixer.scan_core(proj)
        ev[i + 1].record()
    torch.cuda.synchronize()
    t1 = time.time()
for i in range(nblocks):
    print(f"block {i:2d}: {ev[i].elapsed_time(ev[i + 1]) / 20:.3f} ms/step")
pr.terminate()
print("clock samples during the run (t, sm_mhz, W, power_cap, tempC):")
for t, l in rows:
    if t0 - 0.2 <= t <= t1 + 0.2:
        print(f"  {t - t0:6.2f}s  {l}")
