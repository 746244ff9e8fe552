"""Single-token decode step: causal_conv1d_update, selective_state_update and the mixer's cached
branch against the oracle, against the reference's golden decode vectors (oracle/gen_golden.py), and against our own
prefill (decoding token L must equal prefilling L+1 tokens).  Tolerances: 2e-2 (bf16), 1e-4 (fp32).  `pytest -m gpu`."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from oracle import mamba2_ref as R

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


@pytest.fixture(scope="module")
def tv():
    assert torch.cuda.is_available()
    import timeviper_b200
    return timeviper_b200


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("b,dim,K,state_len", [(1, 12288, 4, 4), (2, 576, 4, 6), (3, 40, 2, 2)])
def test_causal_conv1d_update(tv, dtype, b, dim, K, state_len):
    g = torch.Generator(device="cuda").manual_seed(1)
    proj = torch.randn(b, dim + 24, device="cuda", generator=g).to(dtype)
    x = proj[:, 8:8 + dim]                                           # strided rows, as the split of the in_proj output
    state = torch.randn(b, dim, state_len, device="cuda", generator=g).to(dtype)
    w = torch.randn(dim, K, device="cuda", generator=g).to(dtype)
    bias = torch.randn(dim, device="cuda", generator=g).to(dtype)
    ref_out, ref_state = R.causal_conv1d_update_ref(x.cpu(), state.cpu(), w.cpu(), bias.cpu(), "silu")
    out = tv.causal_conv1d_update(x, state, w, bias, "silu")
    assert out.shape == (b, dim) and out.dtype == dtype
    assert relerr(out, ref_out) < TOL[dtype]
    assert torch.equal(state.cpu(), ref_state)                       # a shift: bit exact, in place


@pytest.mark.parametrize("dtype,state_dtype", [(torch.bfloat16, torch.float32), (torch.float32, torch.float32),
                                               (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("with_z", [False, True])
def test_selective_state_update_with_the_reference_s_expanded_views(tv, dtype, state_dtype, with_z):
    b, H, P, G, N = 2, 8, 80, 2, 128
    g = torch.Generator(device="cuda").manual_seed(2)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)     # noqa: E731
    state = rn(b, H, P, N).to(state_dtype)
    x, z = rn(b, H, P).to(dtype), rn(b, H, P).to(dtype)
    dt = rn(b, H).to(dtype)[:, :, None].expand(b, H, P)              # modeling_nano.py:520
    A = (-torch.exp(torch.log(torch.arange(1, H + 1, device="cuda").float())))[:, None, None].expand(H, P, N)   # :514-519
    Bm, Cm = rn(b, G, N).to(dtype), rn(b, G, N).to(dtype)
    D = rn(H).to(dtype)[:, None].expand(H, P)                        # :522
    dt_bias = (rn(H) * 0.5 - 2).to(dtype)[:, None].expand(H, P)      # :521
    kw = dict(D=D, z=z if with_z else None, dt_bias=dt_bias, dt_softplus=True)
    cpu = lambda t: None if t is None else t.cpu()                   # noqa: E731
    ref_out, ref_state = R.selective_state_update_ref(state.cpu(), x.cpu(), dt.cpu(), A.cpu(), Bm.cpu(), Cm.cpu(),
                                                      **{k: (cpu(v) if torch.is_tensor(v) else v) for k, v in kw.items()})
    out = tv.selective_state_update(state, x, dt, A, Bm, Cm, **kw)
    tol = TOL[dtype]
    assert out.shape == (b, H, P) and out.dtype == dtype
    assert relerr(out, ref_out) < tol
    assert relerr(state, ref_state) < (2e-2 if state_dtype == torch.bfloat16 else 1e-5)


def _golden():
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mixer_*.npz")))


def _cache(conv, ssm):
    return types.SimpleNamespace(conv_states=[conv], ssm_states=[ssm], conv_kernel_size=conv.shape[-1])


@pytest.mark.parametrize("path", _golden(), ids=lambda p: os.path.basename(p)[6:-4])
def test_mixer_decode_steps_against_reference_golden(tv, path):
    """Three cached steps of OUR mixer from the reference's prefill cache states vs the reference's own decode outputs
    and final states (fp32)."""
    z = np.load(path)
    hidden, H, P, G, N, Q, L = [int(v) for v in z["dims"]]
    cfg = tv.Mamba2Config(hidden_size=hidden, mamba_num_heads=H, mamba_head_dim=P, n_groups=G, ssm_state_size=N,
                          chunk_size=Q, time_step_limit=tuple(float(v) for v in z["time_step_limit"]))
    mixer = tv.Mamba2MixerPrefill(cfg).cuda()
    keys = ["in_proj.weight", "conv1d.weight", "conv1d.bias", "dt_bias", "A_log", "D", "norm.weight", "out_proj.weight"]
    mixer.load_state_dict({k: torch.from_numpy(z[k]) for k in keys}, strict=True)
    cache = _cache(torch.from_numpy(z["conv_state"]).cuda(), torch.from_numpy(z["ssm_state"]).cuda())
    hs = torch.from_numpy(z["decode_hidden_states"]).cuda()
    outs = []
    with torch.no_grad():
        for i in range(hs.shape[1]):
            outs.append(mixer(hs[:, i:i + 1], cache_params=cache, cache_position=torch.tensor([L + i])))
    assert relerr(torch.cat(outs, dim=1), torch.from_numpy(z["decode_out"])) < 1e-4
    assert relerr(cache.ssm_states[0], torch.from_numpy(z["decode_ssm_state"])) < 1e-4
    # the new columns are in_proj outputs: cuBLAS vs CPU fp32 summation order, not bit-equal
    assert relerr(cache.conv_states[0], torch.from_numpy(z["decode_conv_state"])) < 1e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_decode_continues_our_own_prefill(tv, dtype):
    """prefill(L tokens) + decode(token L) == prefill(L + 1 tokens): last output and both states."""
    torch.manual_seed(3)
    cfg = tv.Mamba2Config(hidden_size=256, mamba_num_heads=16, mamba_head_dim=80, n_groups=2, ssm_state_size=128,
                          chunk_size=128)
    p = R.nemotron_random_params(cfg.hidden_size, 16, 80, 2, 128)
    mixer = tv.Mamba2MixerPrefill(cfg).to(dtype).cuda()
    mixer.load_state_dict({k: v.to(dtype) for k, v in p.items()}, strict=True)
    L = 300
    hs = torch.randn(1, L + 1, cfg.hidden_size, device="cuda").to(dtype)

    def new_cache():
        c = types.SimpleNamespace(conv_states=[None], ssm_states=[None], conv_kernel_size=4)
        c.update_conv_state = lambda layer_idx, new_conv_state, cache_init: c.conv_states.__setitem__(0, new_conv_state.contiguous())
        c.update_ssm_state = lambda layer_idx, new_ssm_state: c.ssm_states.__setitem__(0, new_ssm_state)
        return c

    with torch.no_grad():
        full_cache = new_cache()
        full = mixer(hs, cache_params=full_cache, cache_position=torch.arange(L + 1))
        cache = new_cache()
        mixer(hs[:, :L], cache_params=cache, cache_position=torch.arange(L))
        step = mixer(hs[:, L:], cache_params=cache, cache_position=torch.tensor([L]))
    tol = TOL[dtype]
    assert step.shape == (1, 1, cfg.hidden_size)
    assert relerr(step, full[:, L:]) < tol
    assert relerr(cache.ssm_states[0], full_cache.ssm_states[0]) < tol
    assert relerr(cache.conv_states[0], full_cache.conv_states[0]) < 1e-5 + (tol if dtype == torch.bfloat16 else 0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_decode_step_graph_equals_eager_decode(tv, dtype):
    """The one-launch (CUDA graph) decode step gives bit-identical outputs and states to the eager one over several tokens."""
    torch.manual_seed(5)
    cfg = tv.Mamba2Config(hidden_size=256, mamba_num_heads=16, mamba_head_dim=80, n_groups=2, ssm_state_size=128, chunk_size=128)
    p = R.nemotron_random_params(cfg.hidden_size, 16, 80, 2, 128)
    mixer = tv.Mamba2MixerPrefill(cfg).to(dtype).cuda()
    mixer.load_state_dict({k: v.to(dtype) for k, v in p.items()}, strict=True)
    conv0 = torch.randn(1, cfg.conv_dim, 4, device="cuda").to(dtype)
    ssm0 = torch.randn(1, 16, 80, 128, device="cuda")
    toks = torch.randn(4, 1, 1, cfg.hidden_size, device="cuda").to(dtype)
    eager = types.SimpleNamespace(conv_states=[conv0.clone()], ssm_states=[ssm0.clone()], conv_kernel_size=4)
    graph = types.SimpleNamespace(conv_states=[conv0.clone()], ssm_states=[ssm0.clone()], conv_kernel_size=4)
    with torch.no_grad():
        for t in toks:
            a = mixer.decode_step(t, eager)
            b = mixer.decode_step_graph(t, graph).clone()
            assert torch.equal(a, b)
    assert torch.equal(eager.ssm_states[0], graph.ssm_states[0]) and torch.equal(eager.conv_states[0], graph.conv_states[0])
    assert not torch.equal(graph.ssm_states[0], ssm0)
