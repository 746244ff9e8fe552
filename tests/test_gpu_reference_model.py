"""The UNMODIFIED reference classes on this package's kernels, on the GPU.

`oracle/stage_reference.py` stages the reference's `nano` package (byte for byte, git-ignored) under baseline/_ref/ in
the build container; here `timeviper_b200.patch_reference` rebinds the six operator names of that module
(modeling_nano.py:60-97) and the reference's own `NemotronHMamba2Mixer.cuda_kernels_forward` (:461-668) runs with its
own `HybridMambaAttentionDynamicCache` (:205-360): prefill, three cached decode steps, a padded batch.  Expected values
are the golden vectors the same (unpatched) class produced through its CPU `torch_forward` (tests/golden, G = 1), and
the oracle with the kernel group mapping for G > 1.  fp32: tolerance 1e-4 (north_star).  `pytest -m gpu`."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mamba2_ref as R
from oracle import stage_reference

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not stage_reference.available(), reason="baseline/_ref/nano not staged (run build() "
                                                                         "in the container that has /root/reference)")]
KEYS = ["in_proj.weight", "conv1d.weight", "conv1d.bias", "dt_bias", "A_log", "D", "norm.weight", "out_proj.weight"]


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ref():
    assert torch.cuda.is_available()
    import timeviper_b200 as tv
    mn, Cfg = stage_reference.load()
    tv.patch_reference(mn)
    assert mn.is_fast_path_available and mn.causal_conv1d_fn is tv.causal_conv1d_fn
    return mn, Cfg


def _build(ref, z, batch):
    mn, Cfg = ref
    hidden, H, P, G, N, Q, L = [int(v) for v in z["dims"]]
    lim = tuple(float(v) for v in z["time_step_limit"])
    cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
              mamba_chunk_size=Q, mamba_d_conv=4, mamba_dt_limit=lim, num_hidden_layers=2,
              hybrid_override_pattern="M-", layer_norm_epsilon=1e-5)
    mixer = mn.NemotronHMamba2Mixer(cfg, layer_idx=0).float().eval()
    mixer.load_state_dict({k: torch.from_numpy(z[k]) for k in KEYS}, strict=True)
    mixer = mixer.cuda()
    cache = mn.HybridMambaAttentionDynamicCache(cfg, batch_size=batch, dtype=torch.float32, device="cuda")
    return mixer, cache, (hidden, H, P, G, N, Q, L), lim


def _golden():
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mixer_*.npz")))


@pytest.mark.parametrize("path", _golden(), ids=lambda p: os.path.basename(p)[6:-4])
def test_reference_mixer_prefill_and_decode_on_our_kernels(ref, path):
    z = np.load(path)
    mixer, cache, (hidden, H, P, G, N, Q, L), lim = _build(ref, z, 1)
    hs = torch.from_numpy(z["hidden_states"]).cuda()
    with torch.no_grad():
        out = mixer(hs, cache_params=cache, cache_position=torch.arange(L, device="cuda"))
    sd = {k: torch.from_numpy(z[k]) for k in KEYS}
    if G == 1:
        ref_out, ref_ssm = torch.from_numpy(z["out"]), torch.from_numpy(z["ssm_state"])
    else:   # the golden vector carries torch_forward's h % G mapping (SURVEY finding 4); the fast path maps h // (H/G)
        ref_out, _, ref_ssm = R.mixer_forward_ref(sd, torch.from_numpy(z["hidden_states"]), num_heads=H, head_dim=P,
                                                  n_groups=G, ssm_state_size=N, chunk_size=Q, time_step_limit=lim,
                                                  group_map="kernel")
    assert relerr(out, ref_out) < 1e-4
    assert cache.ssm_states[0].dtype == torch.float32 and relerr(cache.ssm_states[0], ref_ssm) < 1e-4
    assert torch.equal(cache.conv_states[0].cpu(), torch.from_numpy(z["conv_state"]))
    if G != 1 or lim != (0.0, float("inf")):
        return      # decode goldens: G = 1 (mapping) and the default dt limit (the reference's fast decode branch omits the clamp)
    dec_hs = torch.from_numpy(z["decode_hidden_states"]).cuda()
    outs = []
    with torch.no_grad():
        for i in range(dec_hs.shape[1]):
            outs.append(mixer(dec_hs[:, i:i + 1], cache_params=cache, cache_position=torch.tensor([L + i], device="cuda")))
    assert relerr(torch.cat(outs, dim=1), torch.from_numpy(z["decode_out"])) < 1e-4
    assert relerr(cache.ssm_states[0], torch.from_numpy(z["decode_ssm_state"])) < 1e-4
    assert relerr(cache.conv_states[0], torch.from_numpy(z["decode_conv_state"])) < 1e-6


def test_reference_mixer_padded_batch_on_our_kernels(ref):
    """Batch 2, left-padded: the reference's own cuda_kernels_forward applies both masks (:471, :625-627) around our
    conv; expected values from its CPU torch_forward."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "masked_g1_batch2_leftpad37.npz"))
    mixer, cache, (hidden, H, P, G, N, Q, L), lim = _build(ref, z, 2)
    with torch.no_grad():
        out = mixer(torch.from_numpy(z["hidden_states"]).cuda(), cache_params=cache,
                    cache_position=torch.arange(L, device="cuda"),
                    attention_mask=torch.from_numpy(z["attention_mask"]).cuda())
    assert relerr(out, torch.from_numpy(z["out"])) < 1e-4
    assert relerr(cache.ssm_states[0], torch.from_numpy(z["ssm_state"])) < 1e-4
    assert torch.equal(cache.conv_states[0].cpu(), torch.from_numpy(z["conv_state"]))


def test_reference_mixer_bf16_9b_geometry_on_our_kernels(ref):
    """bf16 at the 9B head geometry (tcgen05 path) through the reference's forward, against the oracle fed the same
    bf16-rounded tensors at each kernel boundary (2e-2, north_star)."""
    mn, Cfg = ref
    torch.manual_seed(77)
    hidden, H, P, G, N, Q, L = 256, 16, 80, 2, 128, 128, 700
    cfg = Cfg(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, mamba_n_groups=G, ssm_state_size=N,
              mamba_chunk_size=Q, mamba_d_conv=4, num_hidden_layers=2, hybrid_override_pattern="M-", layer_norm_epsilon=1e-5)
    p = R.nemotron_random_params(hidden, H, P, G, N)
    mixer = mn.NemotronHMamba2Mixer(cfg, layer_idx=0).eval()
    mixer.load_state_dict(p, strict=True)
    mixer = mixer.to(torch.bfloat16).cuda()
    cache = mn.HybridMambaAttentionDynamicCache(cfg, batch_size=1, dtype=torch.bfloat16, device="cuda")
    hs = torch.randn(1, L, hidden).to(torch.bfloat16)
    with torch.no_grad():
        out = mixer(hs.cuda(), cache_params=cache, cache_position=torch.arange(L, device="cuda"))
    pb = {k: v.to(torch.bfloat16).float() for k, v in p.items()}
    ref_out, _, ref_ssm = R.mixer_forward_ref(pb, hs.float(), num_heads=H, head_dim=P, n_groups=G, ssm_state_size=N,
                                              chunk_size=Q, group_map="kernel", round_to=torch.bfloat16)
    assert relerr(out, ref_out) < 2e-2
    assert relerr(cache.ssm_states[0], ref_ssm) < 2e-2
