"""Triangulation against an independent GPU implementation: vLLM's Triton port of the mamba_ssm SSD kernels and gated
RMSNorm (library code in the image, adapted from mamba_ssm v2.2.4 -- the closest available stand-in for the wheels the
reference binds; SURVEY.md 8c "secondary cross-checks").  Not the oracle: the oracle is oracle/mamba2_ref.py, pinned to
the reference's own torch_forward.  Skipped when vLLM's ops cannot be imported.  Tolerance 2e-2 relative (bf16)."""
import pytest
import torch

try:
    from tests.test_gpu_ops import _ssd_inputs, relerr
except ImportError:  # rootdir-relative collection
    from test_gpu_ops import _ssd_inputs, relerr

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def vops():
    assert torch.cuda.is_available()
    try:
        from vllm.model_executor.layers.mamba.ops.layernorm_gated import rms_norm_gated
        from vllm.model_executor.layers.mamba.ops.ssd_combined import mamba_chunk_scan_combined_varlen
    except Exception as e:  # noqa: BLE001  (any import problem of the optional library)
        pytest.skip(f"vLLM Triton ops not importable: {type(e).__name__}: {e}")
    return mamba_chunk_scan_combined_varlen, rms_norm_gated


@pytest.mark.parametrize("L,H,G", [(2048, 16, 2), (1000, 32, 2)])
def test_ssd_and_norm_match_the_triton_port_of_mamba_ssm(vops, L, H, G):
    import timeviper_b200 as tv
    scan_varlen, rms_norm_gated = vops
    P, N, Q = 80, 128, 128
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L, H, P, G, N, torch.bfloat16, seed=41)
    init = torch.randn(1, H, P, N, device="cuda") * 0.5
    nchunks = (L + Q - 1) // Q
    i32 = dict(dtype=torch.int32, device="cuda")
    out_v = torch.empty(L, H, P, dtype=torch.bfloat16, device="cuda")
    st_v = scan_varlen(x[0], dt[0], A, B[0], C[0], Q, torch.tensor([0, L], **i32),
                       torch.tensor(list(range(0, L, Q)) + [L], **i32), torch.tensor([nchunks - 1], **i32),
                       torch.zeros(nchunks, **i32), out_v, D=D.float(), z=None, dt_bias=dt_bias.float(),
                       initial_states=init, dt_softplus=True, state_dtype=torch.float32)
    y, st = tv.mamba_chunk_scan_combined(x, dt, A, B, C, Q, D=D, dt_bias=dt_bias, dt_softplus=True,
                                         initial_states=init, return_final_states=True)
    assert relerr(y[0], out_v) < TOL
    assert relerr(st[0], st_v[0].float()) < TOL
    gate = torch.randn(1, L, H * P + 64, device="cuda").to(torch.bfloat16)[..., :H * P]
    w = (1 + 0.1 * torch.randn(H * P, device="cuda")).to(torch.bfloat16)
    kw = dict(z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False)
    assert relerr(tv.rmsnorm_fn(y.view(1, L, -1), w, None, **kw), rms_norm_gated(y.view(1, L, -1), w, None, **kw)) < TOL
