"""Pure-torch stand-in for mamba_ssm.ops.triton.layernorm_gated.rmsnorm_fn.

TEST INFRASTRUCTURE ONLY.  The reference (modeling_nano.py:73-77) hard-imports this one Triton
function; its semantics (gate-before-norm, grouped RMS, fp32 math) are restated here from the call
site modeling_nano.py:372-380 so that the reference's own torch_forward (modeling_nano.py:671-859)
can run on CPU and produce the golden vectors under tests/golden/.
"""
import torch


def rmsnorm_fn(x, weight, bias=None, z=None, eps=1e-6, group_size=None, norm_before_gate=True,
               upcast=True):
    dtype = x.dtype
    if upcast:
        x = x.float()
        weight = weight.float()
        z = z.float() if z is not None else None
    if z is not None and not norm_before_gate:
        x = x * torch.nn.functional.silu(z)
    d = x.shape[-1]
    g = d if group_size is None else group_size
    xg = x.reshape(*x.shape[:-1], d // g, g)
    rstd = torch.rsqrt(xg.pow(2).mean(-1, keepdim=True) + eps)
    out = (xg * rstd).reshape(x.shape) * weight
    if bias is not None:
        out = out + bias.float()
    if z is not None and norm_before_gate:
        out = out * torch.nn.functional.silu(z)
    return out.to(dtype)
