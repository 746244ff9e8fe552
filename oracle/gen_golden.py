"""Generate tests/golden/mixer_*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Imports /root/reference's ``modeling_nano.py`` (with the one-function pure-torch
``rmsnorm_fn`` shim under oracle/_shim, because the file hard-imports mamba_ssm at :73-77), builds
``NemotronHMamba2Mixer`` (modeling_nano.py:383-885) at small configs, runs its ``forward`` on CPU --
which dispatches to ``torch_forward`` (:862-885 -> :671-859) -- and stores inputs, parameters, the output
and both cache states.  /root/reference does not exist on the GPU box; the committed .npz files do.

    python oracle/gen_golden.py            # rewrites tests/golden/mixer_*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/timeviper/model/llm/llm_repo"

CASES = {
    # name: (hidden, H, P, G, N, Q, L, time_step_limit)
    "g1_ragged300": (96, 4, 80, 1, 128, 128, 300, (0.0, float("inf"))),
    "g1_exact256": (96, 4, 80, 1, 128, 128, 256, (0.0, float("inf"))),
    "g2_literal_short100": (64, 4, 80, 2, 128, 128, 100, (0.0, float("inf"))),
    "g1_dtlimit_q64": (64, 4, 16, 1, 32, 64, 200, (0.01, 0.2)),
}


class _ListOnDevice(list):
    device = torch.device("cpu")


def load_reference():
    sys.path.insert(0, os.path.join(HERE, "_shim"))
    sys.path.insert(0, REF)
    import nano.modeling_nano as mn
    from nano.configuration_nano import NemotronHConfig
    return mn, NemotronHConfig


def main():
    mn, Cfg = load_reference()
    out_dir = os.path.join(HERE, "..", "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (hidden, H, P, G, N, Q, L, lim) in CASES.items():
        torch.manual_seed(1234)
        cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G,
                  ssm_state_size=N, mamba_chunk_size=Q, mamba_d_conv=4, mamba_dt_limit=lim,
                  num_hidden_layers=2, hybrid_override_pattern="M-", layer_norm_epsilon=1e-5)
        mixer = mn.NemotronHMamba2Mixer(cfg, layer_idx=0).float().eval()
        with torch.no_grad():                      # make every term non-degenerate
            mixer.A_log.copy_(torch.log(torch.rand(H) * 15 + 1))
            mixer.dt_bias.copy_(torch.randn(H) * 0.5 - 2.0)
            mixer.D.copy_(torch.randn(H))
            mixer.norm.weight.copy_(1.0 + 0.1 * torch.randn(H * P))
        hs = torch.randn(1, L, hidden)
        cache = mn.HybridMambaAttentionDynamicCache(cfg, batch_size=1, dtype=torch.float32)
        with torch.no_grad():
            out = mixer(hs, cache_params=cache, cache_position=torch.arange(L))
        blob = {k: v.detach().numpy() for k, v in mixer.state_dict().items()}
        prefill_conv, prefill_ssm = cache.conv_states[0].clone(), cache.ssm_states[0].clone()
        # three cached decode steps through the same forward (torch_forward's cache_position > 0 branch, :683-696,
        # :716-773), continuing from the prefill cache
        dec_hs = torch.randn(1, 3, hidden)
        dec_out = []
        # the cached branch reads `cache_params.ssm_states.device` (:718) although the cache keeps per-layer LISTS
        # (:237-254), so as shipped it raises AttributeError; the harness gives the lists a .device instead of
        # touching the reference
        cache.ssm_states, cache.conv_states = _ListOnDevice(cache.ssm_states), _ListOnDevice(cache.conv_states)
        with torch.no_grad():
            for i in range(3):
                dec_out.append(mixer(dec_hs[:, i:i + 1], cache_params=cache, cache_position=torch.tensor([L + i])))
        blob.update(decode_hidden_states=dec_hs.numpy(), decode_out=torch.cat(dec_out, dim=1).numpy(),
                    decode_conv_state=cache.conv_states[0].numpy(), decode_ssm_state=cache.ssm_states[0].numpy())
        blob.update(hidden_states=hs.numpy(), out=out.numpy(),
                    conv_state=prefill_conv.numpy(), ssm_state=prefill_ssm.numpy(),
                    dims=np.array([hidden, H, P, G, N, Q, L], dtype=np.int64),
                    time_step_limit=np.array(lim, dtype=np.float64))
        path = os.path.join(out_dir, f"mixer_{name}.npz")
        np.savez_compressed(path, **blob)
        print(name, "out", tuple(out.shape), "ssm", tuple(cache.ssm_states[0].shape),
              "conv", tuple(cache.conv_states[0].shape), "->", os.path.getsize(path) // 1024, "KiB")


HYBRID_CASES = {
    # name: (hidden, H, P, G, N, Q, attention heads, kv heads, attention head dim, MLP width, pattern, L)
    "m_attn_m_mlp_ragged150": (64, 4, 16, 1, 32, 64, 4, 2, 16, 128, "M*M-", 150),
    "9b_head_geometry_ragged300": (96, 4, 80, 1, 128, 128, 4, 2, 24, 160, "M-M*", 300),   # P=80, N=128, Q=128: tcgen05 path in bf16
}


def main_hybrid():
    """The whole layer loop of the reference's NemotronHModel.forward (modeling_nano.py:1550-1746) on CPU.  The block body
    enters ``torch.cuda.stream(torch.cuda.default_stream(dev))`` (:938); on this CPU-only box the harness replaces those
    two torch functions with no-ops -- the reference stays unmodified."""
    import contextlib
    mn, Cfg = load_reference()
    torch.cuda.default_stream = lambda device=None: None
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    out_dir = os.path.join(HERE, "..", "tests", "golden")
    for name, (hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pattern, L) in HYBRID_CASES.items():
        torch.manual_seed(4321)
        cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
                  mamba_chunk_size=Q, mamba_d_conv=4, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                  layer_norm_epsilon=1e-5, num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd,
                  intermediate_size=mlp, vocab_size=100)
        cfg._attn_implementation = "eager"
        model = mn.NemotronHModel(cfg).float().eval()
        with torch.no_grad():
            for layer in model.layers:
                if layer.block_type == "mamba":
                    layer.mixer.A_log.copy_(torch.log(torch.rand(H) * 15 + 1))
                    layer.mixer.dt_bias.copy_(torch.randn(H) * 0.5 - 2.0)
                    layer.mixer.D.copy_(torch.randn(H))
                layer.norm.weight.copy_(1.0 + 0.1 * torch.randn(hidden))
            x = torch.randn(1, L, hidden)
            out = model(inputs_embeds=x, use_cache=False)
        hs = out[0] if isinstance(out, tuple) else out.last_hidden_state
        blob = {k: v.detach().numpy() for k, v in model.state_dict().items()}
        blob.update(inputs_embeds=x.numpy(), last_hidden_state=hs.numpy(),
                    dims=np.array([hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, L], dtype=np.int64),
                    pattern=np.array(pattern))
        path = os.path.join(out_dir, f"hybrid_{name}.npz")
        np.savez_compressed(path, **blob)
        print(name, "last_hidden_state", tuple(hs.shape), "->", os.path.getsize(path) // 1024, "KiB")



def main_causal_lm():
    """NemotronHForCausalLM.forward (modeling_nano.py:2286-2292, :2414-2433): the hybrid backbone from token ids, then
    ``lm_head(hidden_states).float()`` over the whole sequence.  Stored: the parameters (reference names, ``backbone.*`` +
    ``lm_head.weight``), the token ids, the last hidden state and the fp32 logits of every position."""
    import contextlib
    mn, Cfg = load_reference()
    torch.cuda.default_stream = lambda device=None: None
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    out_dir = os.path.join(HERE, "..", "tests", "golden")
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pattern, L, vocab = 96, 4, 80, 1, 128, 128, 4, 2, 24, 160, "M*M-", 200, 257
    torch.manual_seed(97531)
    cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
              mamba_chunk_size=Q, mamba_d_conv=4, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
              layer_norm_epsilon=1e-5, num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd,
              intermediate_size=mlp, vocab_size=vocab)
    cfg._attn_implementation = "eager"
    model = mn.NemotronHForCausalLM(cfg).float().eval()
    with torch.no_grad():
        for layer in model.backbone.layers:
            if layer.block_type == "mamba":
                layer.mixer.A_log.copy_(torch.log(torch.rand(H) * 15 + 1))
                layer.mixer.dt_bias.copy_(torch.randn(H) * 0.5 - 2.0)
                layer.mixer.D.copy_(torch.randn(H))
            layer.norm.weight.copy_(1.0 + 0.1 * torch.randn(hidden))
        model.lm_head.weight.copy_(torch.randn(vocab, hidden) * 0.2)
        model.backbone.embeddings.weight.copy_(torch.randn(vocab, hidden))
        ids = torch.randint(0, vocab, (1, L))
        out = model(input_ids=ids, use_cache=False, return_dict=True)
        hs = model.backbone(input_ids=ids, use_cache=False)
        hs = hs[0] if isinstance(hs, tuple) else hs.last_hidden_state
    blob = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    blob.update(input_ids=ids.numpy(), logits=out.logits.numpy(), last_hidden_state=hs.numpy(),
                dims=np.array([hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, L, vocab], dtype=np.int64), pattern=np.array(pattern))
    path = os.path.join(out_dir, "causal_lm_MsMd_ids200.npz")
    np.savez_compressed(path, **blob)
    print("causal_lm logits", tuple(out.logits.shape), out.logits.dtype, "->", os.path.getsize(path) // 1024, "KiB")


def main_pdrop(merge_module="no_merge"):
    """TransV / pyramid-drop between layers: NemotronHModel.forward with ``use_pdrop`` (the hooks at
    modeling_nano.py:1634-1689 -> flash_rank_drop :2156 -> pdrop_no_pack :1779-2095), inference, batch 1, ``no_merge`` (the
    default of evaluate.py:167-177): a uniform drop before one layer and attention-ranked drops (query = last prompt token,
    keys = the vision tokens, softmax averaged over heads, top-k) before two attention layers."""
    import contextlib
    mn, Cfg = load_reference()
    torch.cuda.default_stream = lambda device=None: None
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    out_dir = os.path.join(HERE, "..", "tests", "golden")
    hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pattern = 96, 4, 80, 1, 128, 128, 4, 2, 24, 160, "M-M*M-*M"
    pre, V, post = 5, 120, 20
    pdrop_type = "uni_1_0.8-attn_3_0.5-attn_6_0.25"
    torch.manual_seed(8642)
    cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
              mamba_chunk_size=Q, mamba_d_conv=4, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
              layer_norm_epsilon=1e-5, num_attention_heads=ah, num_key_value_heads=kvh, head_dim=ahd,
              intermediate_size=mlp, vocab_size=100, use_pdrop=True, pdrop_type=pdrop_type, merge_module=merge_module)
    cfg._attn_implementation = "eager"
    model = mn.NemotronHModel(cfg).float().eval()
    for k, v in model.pdrop_args.items():          # what NemotronHForCausalLM.set_pdrop_args does (:2459-2462)
        setattr(model, k, v)
    # The reference rebuilds the causal mask after a drop from the OLD cache_position (:1663-1665), which only works when
    # _update_causal_mask returns None, i.e. with flash_attention_2 (:2208-2213) -- the configuration TimeViper runs.
    # flash-attn needs a GPU, so the harness builds the math attention class ("eager", which takes mask None as causal,
    # :1099-1107) and then tells _update_causal_mask that the implementation is flash_attention_2.
    model.config._attn_implementation = "flash_attention_2"
    with torch.no_grad():
        for layer in model.layers:
            if layer.block_type == "mamba":
                layer.mixer.A_log.copy_(torch.log(torch.rand(H) * 15 + 1))
                layer.mixer.dt_bias.copy_(torch.randn(H) * 0.5 - 2.0)
                layer.mixer.D.copy_(torch.randn(H))
            layer.norm.weight.copy_(1.0 + 0.1 * torch.randn(hidden))
        if merge_module == "CrossAttention":       # TransV: alpha is zero at init (no merge); a trained gate is not
            model.alpha.copy_(torch.tensor([0.7, -0.5, 0.9]))
        x = torch.randn(1, pre + V + post, hidden)
        args = dict(is_interleaved=False, first_vision_token_positions=[torch.tensor(pre)], num_vision_tokens=[V],
                    text_prompt_lens=[pre + post])
        out = model(inputs_embeds=x, use_cache=False, return_dict=True, train_pdrop_args=args)
    hs = out.last_hidden_state
    blob = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    blob.update(inputs_embeds=x.numpy(), last_hidden_state=hs.numpy(),
                dims=np.array([hidden, H, P, G, N, Q, ah, kvh, ahd, mlp, pre, V, post], dtype=np.int64),
                pattern=np.array(pattern), pdrop_type=np.array(pdrop_type), merge_module=np.array(merge_module))
    path = os.path.join(out_dir, "pdrop_uni_attn_attn.npz" if merge_module == "no_merge" else "transv_merge_uni_attn_attn.npz")
    np.savez_compressed(path, **blob)
    print("pdrop", tuple(x.shape), "->", tuple(hs.shape), os.path.getsize(path) // 1024, "KiB")

def main_masked():
    """Batch 2, left-padded, with attention_mask: the reference multiplies the padded rows by zero before in_proj (:676)
    and again after the conv (:707; fast path :471 and :625-627), so that silu(conv bias) of a padded position never
    reaches x, B, C."""
    mn, Cfg = load_reference()
    out_dir = os.path.join(HERE, "..", "tests", "golden")
    hidden, H, P, G, N, Q, L = 96, 4, 80, 1, 128, 128, 200
    torch.manual_seed(2468)
    cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
              mamba_chunk_size=Q, mamba_d_conv=4, num_hidden_layers=2, hybrid_override_pattern="M-",
              layer_norm_epsilon=1e-5)
    mixer = mn.NemotronHMamba2Mixer(cfg, layer_idx=0).float().eval()
    with torch.no_grad():
        mixer.A_log.copy_(torch.log(torch.rand(H) * 15 + 1))
        mixer.dt_bias.copy_(torch.randn(H) * 0.5 - 2.0)
        mixer.D.copy_(torch.randn(H))
        mixer.conv1d.bias.copy_(torch.randn(mixer.conv1d.bias.shape) * 0.5)      # a bias that matters at padded rows
    hs = torch.randn(2, L, hidden)
    mask = torch.ones(2, L)
    mask[0, :37] = 0                                   # sequence 0: 37 pad tokens on the left; sequence 1: none
    cache = mn.HybridMambaAttentionDynamicCache(cfg, batch_size=2, dtype=torch.float32)
    with torch.no_grad():
        out = mixer(hs, cache_params=cache, cache_position=torch.arange(L), attention_mask=mask)
    blob = {k: v.detach().numpy() for k, v in mixer.state_dict().items()}
    blob.update(hidden_states=hs.numpy(), attention_mask=mask.numpy(), out=out.numpy(),
                conv_state=cache.conv_states[0].numpy(), ssm_state=cache.ssm_states[0].numpy(),
                dims=np.array([hidden, H, P, G, N, Q, L], dtype=np.int64),
                time_step_limit=np.array((0.0, float("inf")), dtype=np.float64))
    path = os.path.join(out_dir, "masked_g1_batch2_leftpad37.npz")
    np.savez_compressed(path, **blob)
    print("masked batch 2:", tuple(out.shape), "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if "--masked-only" in sys.argv:
        main_masked()
    elif "--causal-lm-only" in sys.argv:
        main_causal_lm()
    elif "--pdrop-only" in sys.argv:
        main_pdrop()
        main_pdrop("CrossAttention")
    else:
        main()
        main_hybrid()
        main_masked()
        main_causal_lm()
