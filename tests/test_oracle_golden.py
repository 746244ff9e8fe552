"""The oracle (oracle/mamba2_ref.py) against golden vectors produced by the UNMODIFIED reference
``NemotronHMamba2Mixer.forward`` -> ``torch_forward`` (modeling_nano.py:671-885); see oracle/gen_golden.py.
CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mamba2_ref as R


def _load(path):
    z = np.load(path)
    d = {k: torch.from_numpy(z[k]) for k in z.files}
    hidden, H, P, G, N, Q, L = [int(v) for v in z["dims"]]
    return d, dict(hidden=hidden, H=H, P=P, G=G, N=N, Q=Q, L=L)


def _cases(golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    return sorted(glob.glob(os.path.join(golden_dir, "mixer_*.npz")))


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p)[6:-4])
def test_mixer_oracle_matches_reference_golden(path):
    d, m = _load(path)
    # the golden vectors come from the literal CPU fallback => its h % G mapping (identical at G=1)
    out, conv_state, ssm_state = R.mixer_forward_ref(
        d, d["hidden_states"], num_heads=m["H"], head_dim=m["P"], n_groups=m["G"],
        ssm_state_size=m["N"], chunk_size=m["Q"], time_step_limit=tuple(d["time_step_limit"].tolist()),
        group_map="torch_forward", dtype=torch.float32)
    assert out.shape == d["out"].shape
    assert relerr(out, d["out"]) < 2e-5
    assert relerr(ssm_state, d["ssm_state"]) < 2e-5
    assert torch.equal(conv_state, d["conv_state"])          # a copy: bit-exact


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p)[6:-4])
def test_sequential_recurrence_matches_golden(path):
    """Second, independent oracle (fp64 token recurrence) against the same reference outputs."""
    d, m = _load(path)
    H, P, G, N = m["H"], m["P"], m["G"], m["N"]
    d_inner, conv_dim = H * P, H * P + 2 * G * N
    proj = torch.nn.functional.linear(d["hidden_states"].double(), d["in_proj.weight"].double())
    gate, xBC, dt = proj.split([d_inner, conv_dim, H], dim=-1)
    xc, _ = R.causal_conv1d_ref(xBC.transpose(1, 2), d["conv1d.weight"].squeeze(1), d["conv1d.bias"],
                                dtype=torch.float64)
    x, Bm, Cm = xc.transpose(1, 2).split([d_inner, G * N, G * N], dim=-1)
    L = m["L"]
    y, h = R.ssd_sequential_ref(x.reshape(1, L, H, P), dt, -torch.exp(d["A_log"].double()),
                                Bm.reshape(1, L, G, N), Cm.reshape(1, L, G, N), D=d["D"],
                                dt_bias=d["dt_bias"], dt_softplus=True,
                                dt_limit=tuple(d["time_step_limit"].tolist()), group_map="torch_forward")
    yn = R.gated_rmsnorm_ref(y.reshape(1, L, d_inner), d["norm.weight"], z=gate, eps=1e-5,
                             group_size=d_inner // G, norm_before_gate=False, dtype=torch.float64)
    out = torch.nn.functional.linear(yn, d["out_proj.weight"].double())
    assert relerr(out, d["out"]) < 2e-5
    assert relerr(h, d["ssm_state"]) < 2e-5


def test_kernel_group_mapping_differs_only_for_multi_group():
    """SURVEY finding 4: h//(H/G) (GPU kernels) vs h%G (torch_forward) agree iff G == 1."""
    torch.manual_seed(0)
    b, L, H, P, N, Q = 1, 70, 4, 8, 16, 32
    x = torch.randn(b, L, H, P); dt = torch.rand(b, L, H) * 0.5; A = -torch.rand(H) - 0.5
    for G, same in ((1, True), (2, False)):
        B = torch.randn(b, L, G, N); C = torch.randn(b, L, G, N)
        y1, s1 = R.ssd_chunked_ref(x, dt, A, B, C, Q, group_map="kernel")
        y2, s2 = R.ssd_chunked_ref(x, dt, A, B, C, Q, group_map="torch_forward")
        assert torch.allclose(y1, y2) == same
        # and the chunked restatement always equals the sequential recurrence for the same mapping
        y3, s3 = R.ssd_sequential_ref(x, dt, A, B, C, group_map="kernel")
        assert relerr(y1, y3) < 1e-5 and relerr(s1, s3) < 1e-5


def test_initial_state_continuation_and_shard_fold():
    """Shard algebra: scanning shards with folded boundary states == one unsharded scan."""
    torch.manual_seed(1)
    b, L, H, P, G, N, Q, W = 1, 256, 4, 8, 2, 16, 32, 4
    x = torch.randn(b, L, H, P, dtype=torch.float64); dt = torch.rand(b, L, H, dtype=torch.float64) * 0.3
    A = -torch.rand(H, dtype=torch.float64) - 0.2
    B = torch.randn(b, L, G, N, dtype=torch.float64); C = torch.randn(b, L, G, N, dtype=torch.float64)
    D = torch.randn(H, dtype=torch.float64)
    y_full, s_full = R.ssd_chunked_ref(x, dt, A, B, C, Q, D=D, dtype=torch.float64)
    sl = [slice(r * L // W, (r + 1) * L // W) for r in range(W)]
    loc = [R.ssd_chunked_ref(x[:, s], dt[:, s], A, B[:, s], C[:, s], Q, D=D, dtype=torch.float64)[1] for s in sl]
    logp = [(dt[:, s] * A).sum(1) for s in sl]
    for r, s in enumerate(sl):
        s_in = R.fold_boundary_states(loc, logp, r)
        y_r, s_out = R.ssd_chunked_ref(x[:, s], dt[:, s], A, B[:, s], C[:, s], Q, D=D,
                                       initial_states=s_in, dtype=torch.float64)
        assert relerr(y_r, y_full[:, s]) < 1e-12
    assert relerr(s_out, s_full) < 1e-12


def test_conv_ref_matches_torch_conv1d_and_halo():
    torch.manual_seed(2)
    b, dim, L, K = 2, 24, 50, 4
    x = torch.randn(b, dim, L); w = torch.randn(dim, K); bias = torch.randn(dim)
    conv = torch.nn.Conv1d(dim, dim, K, groups=dim, padding=K - 1)
    with torch.no_grad():
        conv.weight.copy_(w.unsqueeze(1)); conv.bias.copy_(bias)
        ref = torch.nn.functional.silu(conv(x)[..., :L])             # modeling_nano.py:705
    out, fin = R.causal_conv1d_ref(x, w, bias)
    assert torch.allclose(out, ref, atol=1e-6)
    assert torch.equal(fin, x[..., -(K - 1):])
    # halo: second half continued from the first half's final state == full run
    o1, f1 = R.causal_conv1d_ref(x[..., :20], w, bias)
    o2, _ = R.causal_conv1d_ref(x[..., 20:], w, bias, initial_states=f1)
    assert torch.allclose(torch.cat([o1, o2], -1), out, atol=1e-6)


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p)[6:-4])
def test_decode_step_oracle_matches_reference_golden(path):
    """Three cached single-token steps (torch_forward's cache_position > 0 branch, modeling_nano.py:683-696, 716-773)
    continued from the reference's own prefill cache: outputs and both states after the last step."""
    d, m = _load(path)
    conv, ssm = d["conv_state"], d["ssm_state"]
    outs = []
    for i in range(d["decode_hidden_states"].shape[1]):
        o, conv, ssm = R.mixer_decode_step_ref(
            d, d["decode_hidden_states"][:, i:i + 1], conv, ssm, num_heads=m["H"], head_dim=m["P"], n_groups=m["G"],
            ssm_state_size=m["N"], time_step_limit=tuple(d["time_step_limit"].tolist()))
        outs.append(o)
    assert relerr(torch.cat(outs, dim=1), d["decode_out"]) < 2e-5
    assert relerr(ssm, d["decode_ssm_state"]) < 2e-5
    assert torch.equal(conv, d["decode_conv_state"])


def _hybrid_cases(golden_dir=os.path.join(os.path.dirname(__file__), "golden")):
    return sorted(glob.glob(os.path.join(golden_dir, "hybrid_*.npz")))


@pytest.mark.parametrize("path", _hybrid_cases(), ids=lambda p: os.path.basename(p)[7:-4])
def test_hybrid_stack_oracle_matches_reference_golden(path):
    """Layer loop of the reference's NemotronHModel.forward (Mamba-2 / attention / MLP blocks + norm_f), CPU golden."""
    z = np.load(path)
    d = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("pattern",)}
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, L = [int(v) for v in z["dims"]]
    out = R.hybrid_forward_ref(d, d["inputs_embeds"], pattern=str(z["pattern"]), num_heads=H, head_dim=P, n_groups=G,
                               ssm_state_size=N, chunk_size=Q, attn_heads=ah, kv_heads=kvh, attn_head_dim=ahd,
                               group_map="torch_forward")
    assert relerr(out, d["last_hidden_state"]) < 2e-5


def test_oracle_padding_mask_matches_reference_batch2(golden_dir):
    """Batch 2, left-padded, attention_mask: the oracle applies apply_mask_to_padding_states where torch_forward does
    (modeling_nano.py:676 and :707); golden vector from the reference's own forward (oracle/gen_golden.py::main_masked)."""
    z = np.load(os.path.join(golden_dir, "masked_g1_batch2_leftpad37.npz"))
    hidden, H, P, G, N, Q, L = [int(v) for v in z["dims"]]
    sd = {k: torch.from_numpy(z[k]) for k in ("in_proj.weight", "conv1d.weight", "conv1d.bias", "dt_bias", "A_log", "D",
                                               "norm.weight", "out_proj.weight")}
    hs, mask = torch.from_numpy(z["hidden_states"]), torch.from_numpy(z["attention_mask"])
    out, conv_state, ssm_state = R.mixer_forward_ref(sd, hs, num_heads=H, head_dim=P, n_groups=G, ssm_state_size=N,
                                                     chunk_size=Q, group_map="torch_forward", attention_mask=mask)
    ref = torch.from_numpy(z["out"])
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-5
    ref_ssm = torch.from_numpy(z["ssm_state"])
    assert float((ssm_state - ref_ssm).abs().max() / ref_ssm.abs().max()) < 2e-5
    assert torch.equal(conv_state, torch.from_numpy(z["conv_state"]))
    # without the second mask application the padded rows of sequence 0 leak silu(conv bias) into the state
    out_nomask, _, _ = R.mixer_forward_ref(sd, hs * mask[:, :, None], num_heads=H, head_dim=P, n_groups=G,
                                           ssm_state_size=N, chunk_size=Q, group_map="torch_forward")
    assert float((out_nomask[0] - ref[0]).abs().max() / ref.abs().max()) > 1e-3


def test_causal_lm_logits_oracle_matches_reference_golden():
    """NemotronHForCausalLM.forward of the reference (token ids -> fp32 logits of every position) against the oracle."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "causal_lm_MsMd_ids200.npz"))
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, L, vocab = [int(v) for v in z["dims"]]
    skip = ("pattern", "dims", "input_ids", "logits", "last_hidden_state")
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in skip}
    logits = R.causal_lm_logits_ref(sd, torch.from_numpy(z["input_ids"]), pattern=str(z["pattern"]), num_heads=H, head_dim=P,
                                    n_groups=G, ssm_state_size=N, chunk_size=Q, attn_heads=ah, kv_heads=kvh, attn_head_dim=ahd)
    ref = torch.from_numpy(z["logits"])
    assert logits.shape == ref.shape == (1, L, vocab) and logits.dtype == torch.float32
    assert float((logits - ref).abs().max() / ref.abs().max()) < 2e-5


def test_pdrop_oracle_matches_reference_golden():
    """TransV / pyramid-drop (uniform + two attention-ranked stages) through the reference's own NemotronHModel.forward."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pdrop_uni_attn_attn.npz"))
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pre, V, post = [int(v) for v in z["dims"]]
    skip = ("pattern", "dims", "inputs_embeds", "last_hidden_state", "pdrop_type", "merge_module")
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in skip}
    out = R.hybrid_forward_ref(sd, torch.from_numpy(z["inputs_embeds"]), pattern=str(z["pattern"]), num_heads=H, head_dim=P,
                               n_groups=G, ssm_state_size=N, chunk_size=Q, attn_heads=ah, kv_heads=kvh, attn_head_dim=ahd,
                               pdrop=dict(pdrop_type=str(z["pdrop_type"]), first_vision_token_position=pre,
                                          num_vision_tokens=V, text_prompt_len=pre + post))
    ref = torch.from_numpy(z["last_hidden_state"])
    assert out.shape == ref.shape == (1, pre + int(V * 0.25) + post, hidden)
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-5


def test_transv_merge_oracle_matches_reference_golden():
    """TransV: pyramid-drop WITH the cross-attention merge module (the dropped vision tokens are read by the text tokens
    through a gated cross attention before they go), through the reference's own NemotronHModel.forward."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "transv_merge_uni_attn_attn.npz"))
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pre, V, post = [int(v) for v in z["dims"]]
    skip = ("pattern", "dims", "inputs_embeds", "last_hidden_state", "pdrop_type", "merge_module")
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in skip}
    out = R.hybrid_forward_ref(sd, torch.from_numpy(z["inputs_embeds"]), pattern=str(z["pattern"]), num_heads=H, head_dim=P,
                               n_groups=G, ssm_state_size=N, chunk_size=Q, attn_heads=ah, kv_heads=kvh, attn_head_dim=ahd,
                               pdrop=dict(pdrop_type=str(z["pdrop_type"]), first_vision_token_position=pre,
                                          num_vision_tokens=V, text_prompt_len=pre + post, merge_module="CrossAttention"))
    ref = torch.from_numpy(z["last_hidden_state"])
    assert out.shape == ref.shape and float((out - ref).abs().max() / ref.abs().max()) < 2e-5
    nomerge = np.load(os.path.join(os.path.dirname(__file__), "golden", "pdrop_uni_attn_attn.npz"))["last_hidden_state"]
    assert float(np.abs(nomerge - z["last_hidden_state"]).max()) > 1e-2      # the merge really changes the result


def test_oracle_size_independent_properties():
    """Properties the SSD scan must have whatever the size (the GPU tests use the same ones at full size through
    tools/lin_check.py): exact linearity in x, independence of the chunk size (it only moves rounding points), and
    agreement of the chunked form with the token-by-token recurrence."""
    torch.manual_seed(17)
    b, L, H, P, G, N = 1, 200, 4, 8, 2, 16
    x = torch.randn(b, L, H, P, dtype=torch.float64)
    x2 = torch.randn(b, L, H, P, dtype=torch.float64)
    dt = torch.randn(b, L, H, dtype=torch.float64) * 0.5
    A = -torch.rand(H, dtype=torch.float64) * 3 - 0.05
    B = torch.randn(b, L, G, N, dtype=torch.float64)
    C = torch.randn(b, L, G, N, dtype=torch.float64)
    bias = torch.randn(H, dtype=torch.float64) * 0.5 - 1.0
    kw = dict(dt_bias=bias, dt_softplus=True, dtype=torch.float64)
    y1, s1 = R.ssd_chunked_ref(x, dt, A, B, C, 32, **kw)
    y2, s2 = R.ssd_chunked_ref(x2, dt, A, B, C, 32, **kw)
    y12, s12 = R.ssd_chunked_ref(2.0 * x - 3.0 * x2, dt, A, B, C, 32, **kw)
    assert float((y12 - (2.0 * y1 - 3.0 * y2)).abs().max()) < 1e-10 and float((s12 - (2.0 * s1 - 3.0 * s2)).abs().max()) < 1e-10
    for q in (16, 64, 200, 256):                              # chunk size: same function of the inputs
        yq, sq = R.ssd_chunked_ref(x, dt, A, B, C, q, **kw)
        assert float((yq - y1).abs().max()) < 1e-10 and float((sq - s1).abs().max()) < 1e-10, q
    ys, ss = R.ssd_sequential_ref(x, dt, A, B, C, **kw)
    assert float((ys - y1).abs().max()) < 1e-9 and float((ss - s1).abs().max()) < 1e-9
