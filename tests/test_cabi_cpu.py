"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python operators keep the reference's names/signatures and fail loudly without a GPU."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            names |= set(re.findall(r"\b(tv_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import timeviper_b200._lib as L
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 8
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(declared) == set(L.EXPORTS)
    assert lib.tv_abi_version() == L.TV_ABI_VERSION == 4


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors must have the C layout: compile-free check through field counts and sizes."""
    import timeviper_b200._lib as L
    assert ctypes.sizeof(L.ConvParams) == 6 * 8 + 4 * 4 + 4 * 8 + 2 * 4
    assert ctypes.sizeof(L.RmsnormParams) == 5 * 8 + 8 + 2 * 4 + 3 * 8 + 3 * 4 + 4   # + tail padding
    assert ctypes.sizeof(L.AddRmsnormParams) == 5 * 8 + 8 + 2 * 4 + 4 * 8 + 2 * 4
    assert ctypes.sizeof(L.SsdParams) == 12 * 8 + 7 * 4 + 4 + 15 * 8 + 2 * 4 + 2 * 4 + 4 * 4


def test_decode_struct_layouts():
    import timeviper_b200._lib as L
    assert ctypes.sizeof(L.ConvUpdateParams) == 96      # sizeof(tv_conv1d_update_params), gcc x86-64
    assert ctypes.sizeof(L.SsuParams) == 288            # sizeof(tv_ssu_params)


def test_null_and_invalid_arguments_return_error_codes_without_touching_a_gpu():
    import timeviper_b200._lib as L
    lib = L.load()
    assert lib.tv_causal_conv1d_fwd(None, None) == L.TV_ERR_INVALID
    assert b"null" in lib.tv_last_error()
    p = L.ConvParams(batch=1, dim=12, seqlen=4, width=4, dtype=L.TV_BF16)
    assert lib.tv_causal_conv1d_fwd(ctypes.byref(p), None) == L.TV_ERR_INVALID
    q = L.SsdParams(batch=1, seqlen=8, nheads=3, headdim=8, ngroups=2, dstate=8, chunk_size=64)
    assert lib.tv_ssd_chunk_scan_fwd(ctypes.byref(q), None, 0, None) == L.TV_ERR_INVALID   # 3 % 2 != 0
    assert lib.tv_gated_rmsnorm_fwd(None, None) == L.TV_ERR_INVALID


def test_operator_signatures_match_reference_call_sites():
    import timeviper_b200 as tv
    # causal_conv1d 1.5.x: (x, weight, bias, seq_idx, initial_states, return_final_states, final_states_out, activation)
    assert list(inspect.signature(tv.causal_conv1d_fn).parameters) == [
        "x", "weight", "bias", "seq_idx", "initial_states", "return_final_states", "final_states_out", "activation"]
    # visualize/nano/my_ssd_combined.py:1270-1287
    sig = list(inspect.signature(tv.mamba_chunk_scan_combined).parameters)
    assert sig[:16] == ["x", "dt", "A", "B", "C", "chunk_size", "D", "z", "dt_bias", "initial_states", "seq_idx",
                        "cu_seqlens", "dt_softplus", "dt_limit", "return_final_states", "return_varlen_states"]
    # modeling_nano.py:372-380 keyword call
    assert list(inspect.signature(tv.rmsnorm_fn).parameters) == [
        "x", "weight", "bias", "z", "eps", "group_size", "norm_before_gate", "upcast"]
    # modeling_nano.py:862-869
    assert list(inspect.signature(tv.Mamba2MixerPrefill.forward).parameters) == [
        "self", "hidden_states", "cache_params", "cache_position", "attention_mask", "seq_idx"]


def test_no_cpu_fallback():
    import timeviper_b200 as tv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tv.causal_conv1d_fn(torch.zeros(1, 8, 4), torch.zeros(8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tv.rmsnorm_fn(torch.zeros(2, 8), torch.ones(8), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tv.mamba_chunk_scan_combined(torch.zeros(1, 4, 2, 8), torch.zeros(1, 4, 2), torch.zeros(2),
                                     torch.zeros(1, 4, 1, 8), torch.zeros(1, 4, 1, 8), 64)
    m = tv.Mamba2MixerPrefill(tv.Mamba2Config(hidden_size=32, mamba_num_heads=2, mamba_head_dim=8, n_groups=1,
                                              ssm_state_size=8, chunk_size=64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 32))
    with pytest.raises(NotImplementedError):                       # training fwd+bwd is outside this path
        tv.mamba_split_conv1d_scan_combined()
    with pytest.raises(RuntimeError, match="no CPU fallback"):      # the decode-step operators are CUDA-only too
        tv.causal_conv1d_update(torch.zeros(1, 8), torch.zeros(1, 8, 4), torch.zeros(8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tv.selective_state_update(torch.zeros(1, 2, 4, 8), torch.zeros(1, 2, 4), torch.zeros(1, 2, 4),
                                  torch.zeros(2, 4, 8), torch.zeros(1, 1, 8), torch.zeros(1, 1, 8))


def test_mixer_loads_reference_state_dict(golden_dir):
    """Parameter names/shapes are the reference's (modeling_nano.py:414-451): its state_dict loads strictly."""
    import timeviper_b200 as tv
    z = np.load(os.path.join(golden_dir, "mixer_g1_ragged300.npz"))
    hidden, H, P, G, N, Q, L = [int(v) for v in z["dims"]]
    m = tv.Mamba2MixerPrefill(tv.Mamba2Config(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, n_groups=G,
                                              ssm_state_size=N, chunk_size=Q))
    keys = ["in_proj.weight", "conv1d.weight", "conv1d.bias", "dt_bias", "A_log", "D", "norm.weight", "out_proj.weight"]
    sd = {k: torch.from_numpy(z[k]) for k in keys}
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_patch_reference_rebinds_all_gate_names():
    import types
    import timeviper_b200 as tv
    mod = types.SimpleNamespace(causal_conv1d_fn=None, causal_conv1d_update=None, mamba_chunk_scan_combined=None,
                                mamba_split_conv1d_scan_combined=None, selective_state_update=None, rmsnorm_fn=None,
                                is_fast_path_available=False)
    tv.patch_reference(mod)
    assert mod.is_fast_path_available
    assert all((mod.selective_state_update, mod.mamba_chunk_scan_combined, mod.mamba_split_conv1d_scan_combined,
                mod.causal_conv1d_fn, mod.causal_conv1d_update))                # the gate of modeling_nano.py:89-97
    assert mod.causal_conv1d_fn is tv.causal_conv1d_fn and mod.rmsnorm_fn is tv.rmsnorm_fn


@pytest.mark.skipif(not os.path.isdir("/root/reference/timeviper"), reason="reference tree only exists in the build container")
def test_patch_reference_on_the_real_module():
    """The actual reference module (imported with the rmsnorm shim) takes its fast path after patching."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
    sys.path.insert(0, "/root/reference/timeviper/model/llm/llm_repo")
    import nano.modeling_nano as mn
    import timeviper_b200 as tv
    assert not mn.is_fast_path_available
    saved = {k: getattr(mn, k) for k in ("causal_conv1d_fn", "causal_conv1d_update", "mamba_chunk_scan_combined",
                                         "mamba_split_conv1d_scan_combined", "selective_state_update", "rmsnorm_fn",
                                         "is_fast_path_available")}
    try:
        tv.patch_reference(mn)
        assert mn.is_fast_path_available and mn.mamba_chunk_scan_combined is tv.mamba_chunk_scan_combined
        # the names the prefill branch calls (modeling_nano.py:619, :639, :372) resolve to this package
        src = inspect.getsource(mn.NemotronHMamba2Mixer.cuda_kernels_forward)
        for name in ("causal_conv1d_fn(", "mamba_chunk_scan_combined("):
            assert name in src
    finally:
        for k, v in saved.items():
            setattr(mn, k, v)


def test_streamed_prefill_segment_schedule():
    """Host logic of prefill_from_host: segments tile [0, L) exactly, every boundary but the end is a chunk multiple
    (the SSD state and conv halo are carried across them), no segment exceeds the buffer, and long inputs ramp up /
    down so that the un-overlapped first H2D and last D2H copies are short."""
    import random
    from timeviper_b200.mixer import Mamba2MixerPrefill as M
    rnd = random.Random(0)
    cases = [(131072, 16384), (1000, 256), (300, 128), (5, 128), (128, 128), (70000, 16384)]
    cases += [(rnd.randint(1, 200000), rnd.choice([128, 256, 1024, 4096, 16384])) for _ in range(500)]
    for L, seg in cases:
        b = M._segment_bounds(L, seg, 128)
        sizes = [b1 - b0 for b0, b1 in zip(b, b[1:])]
        assert b[0] == 0 and b[-1] == L and all(n > 0 for n in sizes)
        assert all(x % 128 == 0 for x in b[:-1])
        assert max(sizes) <= seg + 127
    b = M._segment_bounds(131072, 16384, 128)
    assert b[1] == 2048 and b[-1] - b[-2] == 2048


@pytest.mark.skipif(not os.path.isdir("/root/reference/timeviper"), reason="reference tree only exists in the build container")
def test_hybrid_stack_mirrors_the_reference_model_parameters():
    """HybridPrefillStack built from the reference's own NemotronHConfig has exactly the parameter names and shapes of
    the reference NemotronHModel, so its state_dict loads strictly (timeviper_b200/hybrid.py)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
    sys.path.insert(0, "/root/reference/timeviper/model/llm/llm_repo")
    import nano.modeling_nano as mn
    from nano.configuration_nano import NemotronHConfig
    import timeviper_b200 as tv
    rcfg = NemotronHConfig(hidden_size=64, mamba_num_heads=4, mamba_head_dim=16, mamba_n_groups=2, ssm_state_size=32,
                           mamba_chunk_size=64, mamba_d_conv=4, num_hidden_layers=5, hybrid_override_pattern="M*-M-",
                           num_attention_heads=4, num_key_value_heads=2, head_dim=16, intermediate_size=96, vocab_size=50)
    rcfg._attn_implementation = "eager"
    ref = mn.NemotronHModel(rcfg)
    cfg = tv.Mamba2Config.from_hf(rcfg)
    assert (cfg.n_groups, cfg.chunk_size, cfg.intermediate_size_mlp, cfg.hybrid_override_pattern) == (2, 64, 96, "M*-M-")
    ours = tv.HybridPrefillStack(cfg)
    ref_sd, our_sd = ref.state_dict(), ours.state_dict()
    assert set(ref_sd) == set(our_sd)
    assert all(ref_sd[k].shape == our_sd[k].shape for k in ref_sd)
    missing, unexpected = ours.load_state_dict(ref_sd, strict=True)
    assert not missing and not unexpected


def test_hybrid_host_logic_shard_bounds_and_pdrop_schedule():
    """Host-side pieces of the hybrid stack that need no GPU: balanced shard offsets and the pyramid-drop schedule parser
    (modeling_nano.py:1469-1477; default schedule of evaluate.py:167-172)."""
    from timeviper_b200.hybrid import parse_pdrop_type, shard_bounds
    assert shard_bounds(10, 3) == [0, 4, 7, 10] and shard_bounds(8, 2) == [0, 4, 8] and shard_bounds(2, 4) == [0, 1, 2, 2, 2]
    kinds, layers, ratios = parse_pdrop_type("uni_14_0.8-attn_21_0.6-attn_30_0.4-attn_39_0.2")
    assert kinds == ["uni", "attn", "attn", "attn"] and layers == [14, 21, 30, 39] and ratios == [1.0, 0.8, 0.6, 0.4, 0.2]
    import pytest
    with pytest.raises(ValueError):
        parse_pdrop_type("uni_14")
    import timeviper_b200 as tv
    pat = tv.Mamba2Config.nanov2_9b_hybrid().hybrid_override_pattern
    assert len(pat) == 56 and pat.count("*") == 4 and pat.count("M") == 27 and pat.count("-") == 25
    assert [i for i, c in enumerate(pat) if c == "*"] == [14, 21, 30, 39]
