"""BASELINE.json configs[3]: prefill of a random-init 56-layer Nanov2-9B-shaped hybrid stack (27 Mamba-2 / 4 attention /
25 MLP layers; attention at layers 14/21/30/39, SURVEY.md section 8 header) over synthetic video tokens, Mamba layers
on this package's kernels.  Prints tokens/s and the time share of each layer type (CUDA events).
    python tools/run_hybrid_9b.py [tokens] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import timeviper_b200 as tv

L = int(sys.argv[1]) if len(sys.argv) > 1 else 81920            # 5K frames x 16 tokens + text
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
attn = {14, 21, 30, 39}
pat, k = [], 0
for i in range(56):
    if i in attn:
        pat.append("*")
    else:
        pat.append("M-"[k % 2]); k += 1
pat[len(pat) - 1 - pat[::-1].index("-")] = "M"                   # 27 M / 25 -
pattern = "".join(pat)
cfg = tv.Mamba2Config(num_hidden_layers=56, hybrid_override_pattern=pattern, vocab_size=1024)
torch.manual_seed(0)
t0 = time.time()
with torch.device("cuda"):
    model = tv.HybridPrefillStack(cfg).to(torch.bfloat16).eval()
print(f"pattern {pattern}  ({pattern.count('M')} M / {pattern.count('*')} * / {pattern.count('-')} -), "
      f"{sum(p.numel() for p in model.parameters()) / 1e9:.2f} B parameters, built in {time.time() - t0:.1f} s", flush=True)
x = torch.randn(1, L, cfg.hidden_size, device="cuda").to(torch.bfloat16)

# per-layer-type timing through forward hooks (events on the current stream)
ev = []
def pre(m, a):
    e = torch.cuda.Event(enable_timing=True); e.record(); m._e0 = e
def post(m, a, o):
    e = torch.cuda.Event(enable_timing=True); e.record(); ev.append((m.block_type, m._e0, e))
for layer in model.layers:
    layer.register_forward_pre_hook(pre); layer.register_forward_hook(post)

for it in range(iters + 1):
    ev.clear()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = model(inputs_embeds=x)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    share = {}
    for kind, a, b in ev:
        share[kind] = share.get(kind, 0.0) + a.elapsed_time(b)
    tag = "warm-up" if it == 0 else f"iter {it}"
    print(f"{tag}: {L} tokens in {ms:.1f} ms = {L / ms / 1e3:.3f} M tokens/s;  " +
          ", ".join(f"{k} {v:.1f} ms ({100 * v / ms:.0f} %)" for k, v in sorted(share.items())), flush=True)
print("finite:", bool(torch.isfinite(out.float()).all()), " peak memory GB:", round(torch.cuda.max_memory_allocated() / 1e9, 1))
