"""Parity at BASELINE.json's FULL size (Nanov2-9B layer, bf16, batch 1, 131,072 tokens) through properties that do
not need the CPU oracle to process 128K tokens:

  * windows: the oracle recomputes short windows of the full-size GPU result (conv with its 3-row halo, SSD tail
    continued from the GPU state of the prefix, sampled norm rows);
  * linearity: conv (no activation) and the SSD scan are linear in x, and scaling by 2 is exact in bf16/fp32, so
    f(2x) must equal 2 f(x) BIT FOR BIT (subnormal outputs excepted);
  * state passing: scanning the two halves with the carried state equals scanning the whole sequence;
  * causality: changing tokens >= t0 leaves outputs < t0 bit-identical.

Tolerance where a tolerance applies: 2e-2 relative (north_star, bf16).  `pytest -m gpu`."""
import pytest
import torch

from oracle import mamba2_ref as R
try:
    from tests.test_gpu_ops import _ssd_inputs, relerr
except ImportError:  # rootdir-relative collection
    from test_gpu_ops import _ssd_inputs, relerr

pytestmark = pytest.mark.gpu
TOL = 2e-2
L_FULL, H, P, G, N, Q = 131072, 128, 80, 8, 128, 128
CONV_DIM = H * P + 2 * G * N


@pytest.fixture(scope="module")
def tv():
    assert torch.cuda.is_available()
    import timeviper_b200
    return timeviper_b200


def _assert_exact_double(twice, once):
    """twice == 2 * once bit for bit, except where the value is subnormal (flushed to zero by the .ftz math)."""
    bad = twice != once * 2
    assert int(bad.sum()) < 64 and bool((once[bad].float().abs() < 1e-30).all()), int(bad.sum())


@pytest.fixture(scope="module")
def ssd_full(tv):
    x, dt, A, B, C, D, z, dt_bias = _ssd_inputs(1, L_FULL, H, P, G, N, torch.bfloat16, seed=21)
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True)
    out, fin = tv.mamba_chunk_scan_combined(x, dt, A, B, C, Q, **kw)
    torch.cuda.synchronize()
    return dict(x=x, dt=dt, A=A, B=B, C=C, D=D, dt_bias=dt_bias, kw=kw, out=out, fin=fin)


def test_ssd_full_size_linearity_is_bit_exact(tv, ssd_full):
    s = ssd_full
    out2, fin2 = tv.mamba_chunk_scan_combined(s["x"] * 2, s["dt"], s["A"], s["B"], s["C"], Q, **s["kw"])
    _assert_exact_double(out2, s["out"])
    _assert_exact_double(fin2, s["fin"])


def test_ssd_full_size_state_passing_between_halves(tv, ssd_full):
    s = ssd_full
    h = L_FULL // 2
    o1, f1 = tv.mamba_chunk_scan_combined(s["x"][:, :h], s["dt"][:, :h], s["A"], s["B"][:, :h], s["C"][:, :h], Q,
                                          **s["kw"])
    o2, f2 = tv.mamba_chunk_scan_combined(s["x"][:, h:], s["dt"][:, h:], s["A"], s["B"][:, h:], s["C"][:, h:], Q,
                                          initial_states=f1, **s["kw"])
    assert torch.equal(o1, s["out"][:, :h])                      # same chunks, same arithmetic
    assert relerr(o2, s["out"][:, h:]) < 1e-5                    # state crosses HBM in fp32 either way
    assert relerr(f2, s["fin"]) < 1e-5


def test_ssd_full_size_tail_against_oracle(tv, ssd_full):
    """Oracle on the last 256 tokens, continued from the GPU state after the first L-256 tokens."""
    s = ssd_full
    t = L_FULL - 2 * Q
    _, f_prefix = tv.mamba_chunk_scan_combined(s["x"][:, :t], s["dt"][:, :t], s["A"], s["B"][:, :t], s["C"][:, :t], Q,
                                               **s["kw"])
    c = lambda v: v.detach().cpu()
    ref, ref_fin = R.ssd_chunked_ref(c(s["x"][:, t:]), c(s["dt"][:, t:]), c(s["A"]), c(s["B"][:, t:]),
                                     c(s["C"][:, t:]), Q, D=c(s["D"]), dt_bias=c(s["dt_bias"]), dt_softplus=True,
                                     initial_states=c(f_prefix))
    assert relerr(s["out"][:, t:], ref) < TOL
    assert relerr(s["fin"], ref_fin) < TOL


def test_ssd_full_size_causality(tv, ssd_full):
    s = ssd_full
    t0 = 100 * Q + 37
    x2 = s["x"].clone()
    x2[:, t0:] += 1.0
    out2, _ = tv.mamba_chunk_scan_combined(x2, s["dt"], s["A"], s["B"], s["C"], Q, **s["kw"])
    assert torch.equal(out2[:, :t0], s["out"][:, :t0])
    assert not torch.equal(out2[:, t0:t0 + Q], s["out"][:, t0:t0 + Q])


def test_conv_full_size_windows_linearity_causality(tv):
    g = torch.Generator(device="cuda").manual_seed(22)
    proj = torch.randn(1, L_FULL, H * P + CONV_DIM + H, device="cuda", generator=g).to(torch.bfloat16)
    xBC = proj[..., H * P:H * P + CONV_DIM]                       # the strided view the mixer passes
    w = (torch.randn(CONV_DIM, 4, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    b = torch.randn(CONV_DIM, device="cuda", generator=g).to(torch.bfloat16)
    out, fin = tv.causal_conv1d_fn(xBC.transpose(1, 2), w, b, activation="silu", return_final_states=True)
    # windows against the oracle, each with its 3-row halo as initial state
    for t0 in (0, 127, 65536 - 5, L_FULL - 64):
        win = xBC[:, t0:t0 + 64].transpose(1, 2).cpu()
        init = xBC[:, t0 - 3:t0].transpose(1, 2).cpu() if t0 >= 3 else None
        ref, _ = R.causal_conv1d_ref(win, w.cpu(), b.cpu(), init, "silu")
        assert relerr(out[:, :, t0:t0 + 64], ref) < TOL, t0
    assert torch.equal(fin, xBC[:, -3:].transpose(1, 2))
    # linear without bias / activation: bit exact under scaling by 2
    lin = tv.causal_conv1d_fn(xBC.transpose(1, 2), w, None, activation=None)
    lin2 = tv.causal_conv1d_fn((xBC * 2).transpose(1, 2), w, None, activation=None)
    _assert_exact_double(lin2, lin)
    # causal
    t0 = 77777
    x2 = xBC.clone()
    x2[:, t0:] += 1.0
    out2 = tv.causal_conv1d_fn(x2.transpose(1, 2), w, b, activation="silu")
    assert torch.equal(out2[:, :, :t0], out[:, :, :t0])


def test_norm_full_size_sampled_rows(tv):
    g = torch.Generator(device="cuda").manual_seed(23)
    proj = torch.randn(1, L_FULL, H * P + 64, device="cuda", generator=g).to(torch.bfloat16)
    gate = proj[..., :H * P]                                       # strided gate view, as in the mixer
    y = torch.randn(1, L_FULL, H * P, device="cuda", generator=g).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(H * P, device="cuda", generator=g)).to(torch.bfloat16)
    out = tv.rmsnorm_fn(y, w, None, z=gate, eps=1e-5, group_size=H * P // G, norm_before_gate=False)
    rows = torch.cat([torch.tensor([0, 1, L_FULL // 2, L_FULL - 1]),
                      torch.randint(0, L_FULL, (508,), generator=torch.Generator().manual_seed(5))])
    ref = R.gated_rmsnorm_ref(y[0, rows.cuda()].cpu(), w.cpu(), None, gate[0, rows.cuda()].cpu(), 1e-5, H * P // G, False)
    assert relerr(out[0, rows.cuda()], ref) < TOL
    # rows are independent: permuting the rows permutes the result bit for bit
    perm = torch.randperm(L_FULL, device="cuda", generator=g)
    out_p = tv.rmsnorm_fn(y[:, perm], w, None, z=gate[:, perm], eps=1e-5, group_size=H * P // G, norm_before_gate=False)
    assert torch.equal(out_p, out[:, perm])
