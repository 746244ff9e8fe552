"""prefill of the hybrid stack (Mamba-2 layers on our kernels, attention via library SDPA, MLP via
cuBLAS) against the reference's own NemotronHModel.forward (tests/golden/hybrid_*.npz, oracle/gen_golden.py) in fp32,
and against the oracle in bf16.  `pytest -m gpu`."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mamba2_ref as R

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _cases():
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "hybrid_*.npz")))


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p)[7:-4])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_hybrid_prefill_against_reference_golden(path, dtype, tol):
    import timeviper_b200 as tv
    z = np.load(path)
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, L = [int(v) for v in z["dims"]]
    pattern = str(z["pattern"])
    cfg = tv.Mamba2Config(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, n_groups=G, ssm_state_size=N,
                          chunk_size=Q, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                          num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd, intermediate_size_mlp=mlp,
                          vocab_size=100)
    model = tv.HybridPrefillStack(cfg)
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("pattern", "dims", "inputs_embeds", "last_hidden_state")}
    model.load_state_dict(sd, strict=True)                 # the reference NemotronHModel's own parameter names
    model = model.to(dtype).cuda().eval()
    x = torch.from_numpy(z["inputs_embeds"]).to(dtype).cuda()
    out = model(inputs_embeds=x)
    assert out.shape == (1, L, hidden)
    # bf16: several layers of bf16 rounding on a normalised output -- compared with the fp32 reference result
    assert relerr(out, torch.from_numpy(z["last_hidden_state"])) < tol
    ref = R.hybrid_forward_ref(sd, torch.from_numpy(z["inputs_embeds"]), pattern=pattern, num_heads=H, head_dim=P,
                               n_groups=G, ssm_state_size=N, chunk_size=Q, attn_heads=ah, kv_heads=kvh, attn_head_dim=ahd)
    assert relerr(out, ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_causal_lm_last_token_logits_against_reference_golden(dtype, tol):
    """HybridCausalLM (reference parameter names, token ids in) against NemotronHForCausalLM.forward of the reference:
    the last-token fp32 logits it returns by default, and the full tensor on request."""
    import timeviper_b200 as tv
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "causal_lm_MsMd_ids200.npz"))
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, L, vocab = [int(v) for v in z["dims"]]
    pattern = str(z["pattern"])
    cfg = tv.Mamba2Config(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, n_groups=G, ssm_state_size=N,
                          chunk_size=Q, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                          num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd, intermediate_size_mlp=mlp,
                          vocab_size=vocab)
    model = tv.HybridCausalLM(cfg)
    skip = ("pattern", "dims", "input_ids", "logits", "last_hidden_state")
    model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files if k not in skip}, strict=True)
    model = model.to(dtype).cuda().eval()
    ids = torch.from_numpy(z["input_ids"]).cuda()
    ref = torch.from_numpy(z["logits"])
    last = model(input_ids=ids)
    assert last.shape == (1, 1, vocab) and last.dtype == torch.float32
    assert relerr(last, ref[:, -1:]) < tol
    full = model(input_ids=ids, all_positions=True)
    assert full.shape == (1, L, vocab) and relerr(full, ref) < tol
    assert relerr(full[:, -1:], last) < (1e-5 if dtype == torch.float32 else 1e-2)   # same row; cuBLAS picks another kernel for M = 1


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_pyramid_drop_against_reference_golden(dtype, tol):
    """TransV / pyramid-drop between layers (one uniform and two attention-ranked stages) against the
    reference's own NemotronHModel.forward with use_pdrop (tests/golden/pdrop_uni_attn_attn.npz): 145 -> 55 tokens."""
    import timeviper_b200 as tv
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pdrop_uni_attn_attn.npz"))
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pre, V, post = [int(v) for v in z["dims"]]
    pattern = str(z["pattern"])
    cfg = tv.Mamba2Config(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, n_groups=G, ssm_state_size=N,
                          chunk_size=Q, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                          num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd, intermediate_size_mlp=mlp,
                          vocab_size=100)
    model = tv.HybridPrefillStack(cfg)
    skip = ("pattern", "dims", "inputs_embeds", "last_hidden_state", "pdrop_type", "merge_module")
    model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files if k not in skip}, strict=True)
    model = model.to(dtype).cuda().eval()
    x = torch.from_numpy(z["inputs_embeds"]).to(dtype).cuda()
    pd = dict(pdrop_type=str(z["pdrop_type"]), first_vision_token_position=pre, num_vision_tokens=V, text_prompt_len=pre + post)
    out = model(inputs_embeds=x, pdrop=pd)
    ref = torch.from_numpy(z["last_hidden_state"])
    assert out.shape == ref.shape == (1, pre + int(V * 0.25) + post, hidden)
    if dtype == torch.float32:
        assert relerr(out, ref) < tol
    else:
        # bf16 may rank two near-tied vision tokens the other way round, which swaps whole rows: compare the rows that
        # both runs kept (text and prefix rows are always kept) -- and most vision rows must agree
        close = ((out.float().cpu() - ref).abs().amax(-1) / ref.abs().max()) < tol
        assert bool(close[0, :pre].all()) and bool(close[0, -post:].all()) and float(close.float().mean()) > 0.8


def test_transv_merge_module_against_reference_golden():
    """TransV with its cross-attention merge module (reference parameter names ``merge_modules.N.*`` and ``alpha``), fp32,
    against the reference's own forward."""
    import timeviper_b200 as tv
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "transv_merge_uni_attn_attn.npz"))
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pre, V, post = [int(v) for v in z["dims"]]
    pattern, ptype = str(z["pattern"]), str(z["pdrop_type"])
    cfg = tv.Mamba2Config(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, n_groups=G, ssm_state_size=N,
                          chunk_size=Q, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                          num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd, intermediate_size_mlp=mlp,
                          vocab_size=100, merge_module="CrossAttention", pdrop_type=ptype)
    model = tv.HybridPrefillStack(cfg)
    skip = ("pattern", "dims", "inputs_embeds", "last_hidden_state", "pdrop_type", "merge_module")
    model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files if k not in skip}, strict=True)
    model = model.cuda().eval()
    pd = dict(pdrop_type=ptype, first_vision_token_position=pre, num_vision_tokens=V, text_prompt_len=pre + post)
    out = model(inputs_embeds=torch.from_numpy(z["inputs_embeds"]).cuda(), pdrop=pd)
    ref = torch.from_numpy(z["last_hidden_state"])
    assert out.shape == ref.shape and relerr(out, ref) < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rows,d", [(257, 4480), (33, 96), (5, 10240)])
def test_fused_add_rmsnorm_matches_the_eager_norm(dtype, rows, d):
    """csrc/add_rmsnorm.cu against the eager NemotronHRMSNorm (:888-904) on residual + delta (:965): same rounding points
    (sum rounded to the activation dtype, fp32 statistics and weight multiply)."""
    import timeviper_b200 as tv
    if dtype == torch.float32 and d > 5120:
        pytest.skip("fp32 rows of more than 5120 elements are not served by the fused kernel")
    torch.manual_seed(0)
    x = torch.randn(2, rows, d, device="cuda").to(dtype)
    res = (torch.randn(2, rows, d, device="cuda") * 3).to(dtype)
    w = (1 + 0.1 * torch.randn(d, device="cuda")).to(dtype)

    def eager(h):
        h32 = h.to(torch.float32)
        return (w.to(torch.float32) * (h32 * torch.rsqrt(h32.pow(2).mean(-1, keepdim=True) + 1e-5))).to(dtype)
    out, s = tv.ops.add_rmsnorm(x, w, 1e-5, residual=res)
    assert torch.equal(s, res + x)
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-6
    assert relerr(out, eager(res + x)) < tol
    out1, s1 = tv.ops.add_rmsnorm(x, w, 1e-5)
    assert s1 is x and relerr(out1, eager(x)) < tol
    # a strided view (rows of a wider tensor) is taken as it is
    wide = torch.randn(rows, 2 * d, device="cuda").to(dtype)
    assert relerr(tv.ops.add_rmsnorm(wide[:, :d], w, 1e-5)[0], eager(wide[:, :d])) < tol
