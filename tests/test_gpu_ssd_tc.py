"""Parity of the fused tcgen05/TMEM/TMA SSD kernel (bf16, P=80, N=128, Q=128) with the CPU oracle and with
the fp32 CUDA-core kernels.  Tolerance: 2e-2 relative (north_star, bf16).  `pytest -m gpu`."""
import pytest
import torch

from oracle import mamba2_ref as R
try:
    from tests.test_gpu_ops import _cpu, _ssd_inputs, relerr
except ImportError:  # rootdir-relative collection
    from test_gpu_ops import _cpu, _ssd_inputs, relerr

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def tv():
    assert torch.cuda.is_available()
    import timeviper_b200
    assert timeviper_b200.ssd_kernel_family(torch.bfloat16, 80, 128, 128) == "tcgen05", \
        "the tcgen05 kernel family must serve the Nanov2-9B geometry"
    return timeviper_b200


@pytest.mark.parametrize("b,L,H,G", [(1, 128, 4, 1), (1, 5, 2, 2), (1, 1000, 16, 2), (2, 300, 8, 8), (1, 2049, 8, 1)])
def test_tc_matches_oracle(tv, b, L, H, G):
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(b, L, H, 80, G, 128, torch.bfloat16, seed=11)
    init = torch.randn(b, H, 80, 128, device="cuda") * 0.5
    D2 = torch.randn(H, 80, device="cuda").to(torch.bfloat16)          # (nheads, headdim) skip: explicit D*x path
    for kw in (dict(D=D), dict(D=D, z=z, initial_states=init), dict(D=None, dt_limit=(0.01, 0.3)), dict(D=D2),
               dict(D=D2, z=z)):
        out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, dt_bias=dt_bias, dt_softplus=True,
                                                return_final_states=True, **kw)
        torch.cuda.synchronize()
        cx, cdt, cA, cB, cC, cbias = _cpu(x, dt, A, B, C, dt_bias)
        ckw = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in kw.items()}
        ref, ref_fin = R.ssd_chunked_ref(cx, cdt, cA, cB, cC, 128, dt_bias=cbias, dt_softplus=True, **ckw)
        assert out.dtype == torch.bfloat16 and fin.dtype == torch.float32
        assert relerr(out, ref) < TOL, list(kw)
        assert relerr(fin, ref_fin) < TOL, list(kw)


def test_tc_slow_decay_long_memory(tv):
    """Small |A| so that the inter-chunk state carries real weight across many chunks."""
    b, L, H, G = 1, 1536, 8, 2
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(b, L, H, 80, G, 128, torch.bfloat16, seed=12)
    A = A * 0.002
    out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True,
                                            return_final_states=True)
    ref, ref_fin = R.ssd_chunked_ref(*_cpu(x, dt, A, B, C), 128, D=D.cpu(), dt_bias=dt_bias.cpu(), dt_softplus=True)
    assert relerr(out, ref) < TOL and relerr(fin, ref_fin) < TOL


def test_tc_state_summary_matches_oracle(tv):
    b, L, H, G = 1, 640, 8, 2
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(b, L, H, 80, G, 128, torch.bfloat16, seed=13)
    A = A * 0.01
    fin, logp = tv.mamba_chunk_state_summary(x, dt, A, B, 128, dt_bias=dt_bias, dt_softplus=True)
    cx, cdt, cA, cB, cC, cbias = _cpu(x, dt, A, B, C, dt_bias)
    _, ref_fin = R.ssd_chunked_ref(cx, cdt, cA, cB, cC, 128, dt_bias=cbias, dt_softplus=True)
    ref_lp = (R.dt_activate_ref(cdt, cbias, True) * cA).sum(1)
    assert relerr(fin, ref_fin) < TOL and relerr(logp, ref_lp) < 1e-4


def test_tc_matches_oracle_at_9b_dims_16k(tv):
    """BASELINE.json configs[1]: Nanov2-9B layer, bf16, batch 1, seqlen 16K -- the tensor-core path against the CPU oracle
    (the restatement of the reference's torch_forward, oracle/mamba2_ref.py) on identical inputs; ~3 s of CPU time."""
    b, L, H, G = 1, 16384, 128, 8
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(b, L, H, 80, G, 128, torch.bfloat16, seed=14)
    out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True,
                                            return_final_states=True)
    assert tv.ssd_kernel_family(torch.bfloat16, 80, 128, 128) == "tcgen05"
    ref, ref_fin = R.ssd_chunked_ref(*_cpu(x, dt, A, B, C), 128, D=D.cpu(), dt_bias=dt_bias.cpu(), dt_softplus=True)
    assert relerr(out, ref) < TOL and relerr(fin, ref_fin) < TOL
    # the fp32 CUDA-core family on the same inputs (what fp32 callers and other shapes get) agrees as well
    simt, simt_fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, D=D, dt_bias=dt_bias, dt_softplus=True,
                                                  return_final_states=True, _force_simt=True)
    assert relerr(simt, ref) < TOL and relerr(simt_fin, ref_fin) < TOL


def test_reuse_dt_cumsum_refuses_a_workspace_filled_from_other_inputs(tv):
    """`_reuse_dt_cumsum` skips the dt/cumsum pre-kernel and trusts the scratch of the previous call: the library tags
    the scratch with the inputs it was computed from and must refuse a reuse that does not match."""
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, 512, 8, 80, 2, 128, torch.bfloat16, seed=3)
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
    out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, **kw)
    out2, fin2 = tv.mamba_chunk_scan_combined(x, dt, A, B, C, 128, _reuse_dt_cumsum=True, **kw)     # same inputs: fine
    assert torch.equal(out, out2) and torch.equal(fin, fin2)
    with pytest.raises(ValueError, match="reuse_dt_cumsum"):                                         # another dt tensor
        tv.mamba_chunk_scan_combined(x, dt.clone(), A, B, C, 128, _reuse_dt_cumsum=True, **kw)
    with pytest.raises(ValueError, match="reuse_dt_cumsum"):                                         # other dims
        tv.mamba_chunk_scan_combined(x[:, :256], dt[:, :256], A, B[:, :256], C[:, :256], 128, _reuse_dt_cumsum=True, **kw)


def test_chunk_size_256_runs_on_the_tensor_core_kernel(tv):
    """The reference class default is chunk_size = 256 (configuration_nano.py:137-175): any multiple of 128 is served by the
    tcgen05 kernel (which walks 128-token chunks internally -- the chunk size only moves rounding points)."""
    assert tv.ssd_kernel_family(torch.bfloat16, 80, 128, 256) == "tcgen05"
    b, L, H, G = 1, 1500, 8, 2
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(b, L, H, 80, G, 128, torch.bfloat16, seed=21)
    init = torch.randn(b, H, 80, 128, device="cuda") * 0.5
    for q in (256, 512):
        out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, q, D=D, dt_bias=dt_bias, dt_softplus=True,
                                                initial_states=init, return_final_states=True)
        ref, ref_fin = R.ssd_chunked_ref(*_cpu(x, dt, A, B, C), q, D=D.cpu(), dt_bias=dt_bias.cpu(), dt_softplus=True,
                                         initial_states=init.cpu())
        assert relerr(out, ref) < TOL and relerr(fin, ref_fin) < TOL, q
