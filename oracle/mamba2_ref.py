"""CPU oracle for the Mamba-2 mixer prefill path of xiaomi-research/timeviper.

TEST INFRASTRUCTURE -- NOT THE PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The product
(``timeviper_b200``) never does; it raises when its CUDA extension is missing.

What is restated here (all paths relative to /root/reference):

* ``NemotronHMamba2Mixer.torch_forward``   timeviper/model/llm/llm_repo/nano/modeling_nano.py:671-859
  - conv:  ``act(conv1d(x^T)[..., :L])``                                        :705   (conv def :414-421)
  - dt:    ``clamp(softplus(dt + dt_bias), *time_step_limit)``                    :776-777
  - SSD:   chunked state-space-duality scan                                        :778-847
  - cache: conv state = last ``conv_kernel`` *pre-conv* columns, left zero-padded  :698-703
           ssm state  = state after the last real token, fp32 (b, H, P, N)         :829, :850-851
  - norm:  ``rmsnorm_fn(y, w, z=gate, group_size=d_inner/G, norm_before_gate=False)``  :363-380, :853
* GPU-path stage boundaries (``mamba_ssm`` kernels called through the vendored wrapper)
  visualize/nano/my_ssd_combined.py:743-843, ``dt_bias_activate`` :605-615.

Deliberate deviation (finding 4): torch_forward expands B/C groups to heads with
``repeat`` (:781-782, head h -> group h % G).  The GPU kernels the reference actually ships results from
(and the reference's own fixed copy visualize/nano/modeling_nano.py:997-998) use head h -> group
h // (H/G).  ``group_map="kernel"`` (default) is that mapping; ``group_map="torch_forward"`` reproduces
the literal CPU fallback, and is what the golden vectors generated from the unmodified reference are
checked with.

Parity pinning: the reference has no tests or golden vectors for this path.  The oracle is pinned
against outputs of the reference's own ``torch_forward`` run in the build container
(oracle/gen_golden.py -> tests/golden/*.npz).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# causal depthwise conv1d + activation                    (modeling_nano.py:414-421, :619-624, :705)
# --------------------------------------------------------------------------------------------
def causal_conv1d_ref(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      initial_states: Optional[torch.Tensor] = None, activation: Optional[str] = "silu",
                      dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """x (b, dim, L), weight (dim, K), bias (dim,), initial_states (b, dim, K-1) = the K-1 columns that
    precede x.  Returns (out (b, dim, L), final_states (b, dim, K-1)) in ``dtype``."""
    b, dim, L = x.shape
    K = weight.shape[1]
    xf = x.to(dtype)
    if initial_states is None:
        left = torch.zeros(b, dim, K - 1, dtype=dtype)
    else:
        left = initial_states.to(dtype)
    xp = torch.cat([left, xf], dim=-1)                       # zero left pad == Conv1d(padding=K-1)[..., :L]
    out = F.conv1d(xp, weight.to(dtype).unsqueeze(1), None if bias is None else bias.to(dtype),
                   groups=dim)
    if activation in ("silu", "swish"):
        out = F.silu(out)
    elif activation is not None:
        raise ValueError(activation)
    return out, xp[..., -(K - 1):].contiguous()


def conv_cache_state_ref(xBC: torch.Tensor, conv_kernel: int) -> torch.Tensor:
    """modeling_nano.py:596-610: (b, L, conv_dim) pre-conv activations -> (b, conv_dim, conv_kernel)."""
    xt = xBC.transpose(1, 2)
    return F.pad(xt, (conv_kernel - xt.shape[-1], 0))


# --------------------------------------------------------------------------------------------
# dt activation                      (modeling_nano.py:776-777; my_ssd_combined.py:605-615)
# --------------------------------------------------------------------------------------------
def dt_activate_ref(dt, dt_bias=None, dt_softplus=False, dt_limit=(0.0, float("inf")),
                    dtype=torch.float32):
    dt = dt.to(dtype)
    if dt_bias is not None:
        dt = dt + dt_bias.to(dtype)
    if dt_softplus:
        dt = torch.where(dt <= 20.0, F.softplus(dt), dt)
    lo, hi = float(dt_limit[0]), float(dt_limit[1])
    return dt.clamp(min=lo, max=hi)


def _expand_groups(t: torch.Tensor, nheads: int, group_map: str) -> torch.Tensor:
    """(b, L, G, N) -> (b, L, H, N)."""
    G = t.shape[2]
    if group_map == "kernel":            # h -> h // (H/G)   (mamba_ssm kernels; the shipped GPU path)
        return t.repeat_interleave(nheads // G, dim=2)
    if group_map == "torch_forward":     # h -> h % G        (modeling_nano.py:781-782, literal)
        return t.repeat(1, 1, nheads // G, 1)
    raise ValueError(group_map)


# --------------------------------------------------------------------------------------------
# SSD, chunked (memory-lean restatement of torch_forward :778-847)
# --------------------------------------------------------------------------------------------
def ssd_chunked_ref(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None,
                    dt_softplus=False, dt_limit=(0.0, float("inf")), group_map="kernel",
                    dtype=torch.float32):
    """Same contract as mamba_chunk_scan_combined (my_ssd_combined.py:1270-1306).

    x (b,L,H,P)  dt (b,L,H)  A (H,)  B,C (b,L,G,N)  D (H,)|(H,P)  z (b,L,H,P)  initial_states (b,H,P,N)
    Returns (y (b,L,H,P), final_states (b,H,P,N)), both in ``dtype``.

    Follows torch_forward step by step but one chunk at a time, so that the (b,c,l,s,h,n) tensor of
    :803 is never materialised (34 GB at the 9B dims)."""
    b, L, H, P = x.shape
    N = B.shape[-1]
    Q = int(chunk_size)
    xf = x.to(dtype)
    dtf = dt_activate_ref(dt, dt_bias, dt_softplus, dt_limit, dtype)          # :776-777
    Af = A.to(dtype)
    Bh = _expand_groups(B.to(dtype), H, group_map)                            # :779-782
    Ch = _expand_groups(C.to(dtype), H, group_map)
    y = torch.empty(b, L, H, P, dtype=dtype)
    state = (torch.zeros(b, H, P, N, dtype=dtype) if initial_states is None
             else initial_states.to(dtype).clone())                           # :821-824
    for s in range(0, L, Q):
        e = min(s + Q, L)                      # ragged tail == zero padding of :783 (dt=0, x=0)
        xc, dtc, Bc, Cc = xf[:, s:e], dtf[:, s:e], Bh[:, s:e], Ch[:, s:e]
        dA = dtc * Af                                                          # :789   (b,q,H)
        cs = torch.cumsum(dA, dim=1)                                           # :796
        # L = exp(segsum(A)) :800   (lower-triangular incl. diagonal)
        seg = cs[:, :, None, :] - cs[:, None, :, :]                            # (b,m,k,H)
        q = e - s
        tri = torch.tril(torch.ones(q, q, dtype=torch.bool))
        Lm = torch.exp(seg.masked_fill(~tri[None, :, :, None], -float("inf")))
        G_ = torch.einsum("bmhn,bkhn->bmkh", Cc, Bc)                           # :803-804
        M_ = G_ * Lm                                                           # :807-808
        xdt = xc * dtc[..., None]                                              # :788
        y_diag = torch.einsum("bmkh,bkhp->bmhp", M_, xdt)                      # :811
        # state -> output (:833-836) uses the state *entering* the chunk
        y_off = torch.einsum("bmhn,bhpn->bmhp", Cc, state) * torch.exp(cs)[..., None]
        y[:, s:e] = y_diag + y_off
        # chunk state (:815-817) and inter-chunk recurrence (:826-829)
        decay = torch.exp(cs[:, -1:, :] - cs)                                  # (b,q,H)
        new = torch.einsum("bkhn,bkhp->bhpn", Bc * decay[..., None], xdt)
        state = state * torch.exp(cs[:, -1, :])[:, :, None, None] + new
    if D is not None:                                                          # :785, :843
        Df = D.to(dtype)
        y = y + xf * (Df[None, None, :, None] if Df.dim() == 1 else Df[None, None])
    if z is not None:                  # ssd_chunk_scan epilogue: out *= silu(z) after the D skip
        y = y * F.silu(z.to(dtype))
    return y, state


# --------------------------------------------------------------------------------------------
# SSD, token-sequential recurrence -- an independent second oracle
# --------------------------------------------------------------------------------------------
def ssd_sequential_ref(x, dt, A, B, C, D=None, z=None, dt_bias=None, initial_states=None,
                       dt_softplus=False, dt_limit=(0.0, float("inf")), group_map="kernel",
                       dtype=torch.float64):
    """h_t = exp(dt_t A) h_{t-1} + dt_t x_t (x) B_t ;  y_t = C_t . h_t + D x_t."""
    b, L, H, P = x.shape
    N = B.shape[-1]
    xf = x.to(dtype)
    dtf = dt_activate_ref(dt, dt_bias, dt_softplus, dt_limit, dtype)
    Af = A.to(dtype)
    Bh = _expand_groups(B.to(dtype), H, group_map)
    Ch = _expand_groups(C.to(dtype), H, group_map)
    h = (torch.zeros(b, H, P, N, dtype=dtype) if initial_states is None
         else initial_states.to(dtype).clone())
    y = torch.empty(b, L, H, P, dtype=dtype)
    for t in range(L):
        a = torch.exp(dtf[:, t] * Af)                                          # (b,H)
        h = h * a[:, :, None, None] + (dtf[:, t, :, None] * xf[:, t])[..., None] * Bh[:, t, :, None, :]
        y[:, t] = torch.einsum("bhpn,bhn->bhp", h, Ch[:, t])
    if D is not None:
        Df = D.to(dtype)
        y = y + xf * (Df[None, None, :, None] if Df.dim() == 1 else Df[None, None])
    if z is not None:
        y = y * F.silu(z.to(dtype))
    return y, h


# --------------------------------------------------------------------------------------------
# gated grouped RMSNorm                                      (modeling_nano.py:363-380, :853)
# --------------------------------------------------------------------------------------------
def gated_rmsnorm_ref(x, weight, bias=None, z=None, eps=1e-6, group_size=None, norm_before_gate=True,
                      dtype=torch.float32):
    xf = x.to(dtype)
    if z is not None and not norm_before_gate:
        xf = xf * F.silu(z.to(dtype))
    d = xf.shape[-1]
    g = d if group_size is None else int(group_size)
    xg = xf.reshape(*xf.shape[:-1], d // g, g)
    out = (xg * torch.rsqrt(xg.pow(2).mean(-1, keepdim=True) + eps)).reshape(xf.shape) * weight.to(dtype)
    if bias is not None:
        out = out + bias.to(dtype)
    if z is not None and norm_before_gate:
        out = out * F.silu(z.to(dtype))
    return out


# --------------------------------------------------------------------------------------------
# the whole mixer (torch_forward :671-859), from a plain dict of parameters
# --------------------------------------------------------------------------------------------
def mixer_forward_ref(p: dict, hidden_states: torch.Tensor, *, num_heads: int, head_dim: int,
                      n_groups: int, ssm_state_size: int, chunk_size: int, conv_kernel: int = 4,
                      eps: float = 1e-5, time_step_limit: Sequence[float] = (0.0, float("inf")),
                      group_map: str = "kernel", dtype=torch.float32, round_to=None, attention_mask=None):
    """p: in_proj.weight (W,hidden), conv1d.weight (conv_dim,1,K), conv1d.bias, dt_bias, A_log, D,
    norm.weight, out_proj.weight.  Returns (out (b,L,hidden), conv_state (b,conv_dim,K), ssm_state (b,H,P,N)).

    ``round_to`` (e.g. torch.bfloat16) re-rounds the tensors that cross a kernel boundary in the GPU path
    (in_proj output, conv output, scan output, norm output) so that a bf16 product run is compared with
    an oracle fed the *same* rounded intermediates."""
    rnd = (lambda t: t) if round_to is None else (lambda t: t.to(round_to).to(dtype))
    H, P, G, N = num_heads, head_dim, n_groups, ssm_state_size
    d_inner = H * P
    conv_dim = d_inner + 2 * G * N
    b, L, _ = hidden_states.shape
    # apply_mask_to_padding_states (:189-201): only for batch > 1 and L > 1; before in_proj (:676) and after the conv (:707)
    masked = attention_mask is not None and attention_mask.shape[0] > 1 and attention_mask.shape[1] > 1
    if masked:
        hidden_states = hidden_states.to(dtype) * attention_mask[:, :, None].to(dtype)
    proj = rnd(F.linear(hidden_states.to(dtype), p["in_proj.weight"].to(dtype)))          # :677
    gate, xBC, dt = proj.split([d_inner, conv_dim, H], dim=-1)                            # :679-681
    conv_state = conv_cache_state_ref(xBC, conv_kernel)                                    # :698-703
    xBC_c, _ = causal_conv1d_ref(xBC.transpose(1, 2), p["conv1d.weight"].squeeze(1),
                                 p.get("conv1d.bias"), activation="silu", dtype=dtype)    # :705
    xBC_c = rnd(xBC_c.transpose(1, 2))
    if masked:
        xBC_c = xBC_c * attention_mask[:, :, None].to(dtype)                              # :707
    x, Bm, Cm = xBC_c.split([d_inner, G * N, G * N], dim=-1)                               # :708-712
    A = -torch.exp(p["A_log"].float())                                                     # :715
    y, ssm_state = ssd_chunked_ref(x.reshape(b, L, H, P), dt, A, Bm.reshape(b, L, G, N),
                                   Cm.reshape(b, L, G, N), chunk_size, D=p["D"], z=None,
                                   dt_bias=p["dt_bias"], dt_softplus=True, dt_limit=time_step_limit,
                                   group_map=group_map, dtype=dtype)                       # :776-851
    y = rnd(y.reshape(b, L, d_inner))
    yn = rnd(gated_rmsnorm_ref(y, p["norm.weight"], None, z=gate, eps=eps, group_size=d_inner // G,
                               norm_before_gate=False, dtype=dtype))                       # :853
    out = F.linear(yn, p["out_proj.weight"].to(dtype))                                     # :858
    return out, conv_state, ssm_state


# --------------------------------------------------------------------------------------------
# single-token decode step: torch_forward's cached branch, modeling_nano.py:683-696, 716-773;
# operator contracts of the fast path at :495-539 (causal_conv1d_update, selective_state_update)
# --------------------------------------------------------------------------------------------
def causal_conv1d_update_ref(x: torch.Tensor, conv_state: torch.Tensor, weight: torch.Tensor,
                             bias: Optional[torch.Tensor] = None, activation: Optional[str] = None,
                             dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """x (b, dim) new pre-conv column, conv_state (b, dim, state_len >= K) -> (out (b, dim), new conv_state).
    The state is shifted left by one column and x appended (:338-344 with cache_init=False); the output is the dot
    product of its last K columns with the weights (:689-696)."""
    K = weight.shape[1]
    new_state = torch.cat([conv_state[..., 1:], x.to(conv_state.dtype)[..., None]], dim=-1)
    out = (new_state[..., -K:].to(dtype) * weight.to(dtype)).sum(-1)
    if bias is not None:
        out = out + bias.to(dtype)
    if activation in ("silu", "swish"):
        out = F.silu(out)
    elif activation is not None:
        raise ValueError(activation)
    return out, new_state


def selective_state_update_ref(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False,
                               dtype: torch.dtype = torch.float32):
    """state (b,H,P,N), x (b,H,P), dt (b,H,P), A (H,P,N), B/C (b,G,N), D (H,P), z (b,H,P), dt_bias (H,P)
    -> (out (b,H,P), new state (b,H,P,N) in ``dtype``).  :716-768; head h reads group h // (H/G) (:740-742)."""
    b, H, P, N = state.shape
    G = B.shape[1]
    dt = dt.to(dtype)
    if dt_bias is not None:
        dt = dt + dt_bias.to(dtype)
    if dt_softplus:
        dt = torch.where(dt <= 20.0, F.softplus(dt), dt)
    dA = torch.exp(dt[..., None] * A.to(dtype))                                   # (b,H,P,N)
    Bh = B.to(dtype).repeat_interleave(H // G, dim=1)                              # (b,H,N)
    Ch = C.to(dtype).repeat_interleave(H // G, dim=1)
    new_state = state.to(dtype) * dA + (dt[..., None] * Bh[:, :, None, :]) * x.to(dtype)[..., None]
    out = (new_state * Ch[:, :, None, :]).sum(-1)
    if D is not None:
        out = out + x.to(dtype) * D.to(dtype)
    if z is not None:
        out = out * F.silu(z.to(dtype))
    return out, new_state


def mixer_decode_step_ref(p: dict, hidden_states: torch.Tensor, conv_state: torch.Tensor, ssm_state: torch.Tensor, *,
                          num_heads: int, head_dim: int, n_groups: int, ssm_state_size: int, eps: float = 1e-5,
                          time_step_limit: Sequence[float] = (0.0, float("inf")), dtype=torch.float32):
    """One cached decode step of the mixer: hidden_states (b,1,hidden), conv_state (b,conv_dim,K), ssm_state
    (b,H,P,N) -> (out (b,1,hidden), new conv_state, new ssm_state)."""
    H, P, G, N = num_heads, head_dim, n_groups, ssm_state_size
    d_inner, conv_dim = H * P, H * P + 2 * G * N
    b = hidden_states.shape[0]
    proj = F.linear(hidden_states.to(dtype), p["in_proj.weight"].to(dtype))[:, 0]         # :677
    gate, xBC, dt = proj.split([d_inner, conv_dim, H], dim=-1)
    xBC_c, conv_state = causal_conv1d_update_ref(xBC, conv_state, p["conv1d.weight"].squeeze(1),
                                                 p.get("conv1d.bias"), "silu", dtype)      # :684-696
    x, Bm, Cm = xBC_c.split([d_inner, G * N, G * N], dim=-1)
    A = -torch.exp(p["A_log"].float())[:, None, None].expand(H, P, N)                      # :715, :727
    dtv = dt[:, :, None].expand(b, H, P)
    dtv = torch.clamp(F.softplus(dtv + p["dt_bias"].to(dtype)[:, None]), time_step_limit[0], time_step_limit[1])
    y, ssm_state = selective_state_update_ref(ssm_state, x.reshape(b, H, P), dtv, A, Bm.reshape(b, G, N),
                                              Cm.reshape(b, G, N), D=p["D"].to(dtype)[:, None].expand(H, P),
                                              dtype=dtype)                                 # :720-768
    yn = gated_rmsnorm_ref(y.reshape(b, 1, d_inner), p["norm.weight"], None, z=gate[:, None], eps=eps,
                           group_size=d_inner // G, norm_before_gate=False, dtype=dtype)   # :853
    return F.linear(yn, p["out_proj.weight"].to(dtype)), conv_state, ssm_state             # :858


# --------------------------------------------------------------------------------------------
# hybrid stack prefill: NemotronHModel layer loop, modeling_nano.py:1550-1746; block :906-967;
# attention :1012-1117 (GQA, no rotary embedding, causal); MLP :970-996 (squared ReLU); RMSNorm :888-904
# --------------------------------------------------------------------------------------------
def rmsnorm_ref(x, weight, eps, dtype=torch.float32):
    h = x.to(dtype)
    return weight.to(dtype) * (h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps))


def parse_pdrop_type(pdrop_type: str):
    """'type_layer_ratio-...' -> (types, layers, ratios with a leading 1)   (modeling_nano.py:1469-1477)"""
    parts = [t.split("_") for t in pdrop_type.split("-")]
    return [t[0] for t in parts], [int(t[1]) for t in parts], [1.0] + [float(t[2]) for t in parts]


def pdrop_select_ref(h, stage, kind, ratios, wq, wk, attn_heads, kv_heads, attn_head_dim, vision_index, num_vision_tokens,
                     text_prompt_len):
    """pdrop_no_pack for inference, one sample (modeling_nano.py:1779-1988): the sorted sequence indices of the vision
    tokens that survive stage `stage`, and the index of the first token after the vision block.  h: (L, hidden)."""
    image_tokens = int(num_vision_tokens * ratios[stage])                                       # :1795-1802
    keep = int(num_vision_tokens * ratios[stage + 1])
    if "attn" in kind:
        L = h.shape[0]
        q = F.linear(h, wq).view(L, attn_heads, attn_head_dim).transpose(0, 1)                   # :1833-1845 (raw features)
        k = F.linear(h, wk).view(L, kv_heads, attn_head_dim).transpose(0, 1)
        k = k.repeat_interleave(attn_heads // kv_heads, dim=0)
        pq = text_prompt_len + image_tokens - 1                                                  # :1917-1924: last prompt token
        w = (q[:, pq:pq + 1] @ k.transpose(1, 2)) / math.sqrt(attn_head_dim)                     # (heads, 1, L)
        mask = torch.zeros(L, dtype=h.dtype)
        mask[pq + 1:] = float("-inf")                                                            # the causal row of the query
        w = F.softmax(w + mask, dim=-1, dtype=torch.float32).to(h.dtype)                         # :1932-1937
        w = w.mean(0)[:, vision_index:vision_index + image_tokens].mean(0)                       # :1939-1943
        top = w.topk(keep).indices
    elif "uni" in kind:
        top = torch.linspace(0, image_tokens - 1, keep, dtype=torch.long)                        # :1950-1957
    else:
        raise NotImplementedError(kind)
    return (top + vision_index).sort().values, vision_index + image_tokens                      # :1961-1965


def hybrid_forward_ref(p: dict, inputs_embeds: torch.Tensor, *, pattern: str, num_heads: int, head_dim: int,
                       n_groups: int, ssm_state_size: int, chunk_size: int, attn_heads: int, kv_heads: int,
                       attn_head_dim: int, eps: float = 1e-5, group_map: str = "kernel", dtype=torch.float32,
                       pdrop: dict = None):
    """p: a NemotronHModel state_dict (layers.N.norm.weight, layers.N.mixer.*, norm_f.weight).  Returns the last hidden
    states (b, L, hidden) after norm_f.  pdrop = dict(pdrop_type, first_vision_token_position, num_vision_tokens,
    text_prompt_len): TransV / pyramid-drop before the listed layers (batch 1, no merge module)."""
    h = inputs_embeds.to(dtype)
    b, L, _ = h.shape
    if pdrop is not None:
        assert b == 1
        kinds, layers, ratios = parse_pdrop_type(pdrop["pdrop_type"])
    for i, kind in enumerate(pattern):
        pre = f"layers.{i}."
        if pdrop is not None and i in layers:                                                      # :1634-1666
            st = layers.index(i)
            wq = p.get(pre + "mixer.q_proj.weight"); wk = p.get(pre + "mixer.k_proj.weight")
            top, start = pdrop_select_ref(h[0], st, kinds[st], ratios, None if wq is None else wq.to(dtype),
                                          None if wk is None else wk.to(dtype), attn_heads, kv_heads, attn_head_dim,
                                          pdrop["first_vision_token_position"], pdrop["num_vision_tokens"], pdrop["text_prompt_len"])
            vi = pdrop["first_vision_token_position"]
            text = h[:, start:]
            if pdrop.get("merge_module", "no_merge") == "CrossAttention":                          # TransV, :1748-1777
                keep_mask = torch.ones(start - vi, dtype=torch.bool)
                keep_mask[top - vi] = False
                dropped = h[:, vi:start][:, keep_mask]                                             # the tokens about to go
                mp_ = f"merge_modules.{st}."
                q = F.linear(text, p[mp_ + "q_proj.weight"].to(dtype)).view(1, -1, attn_heads, attn_head_dim).transpose(1, 2)
                k = F.linear(dropped, p[mp_ + "k_proj.weight"].to(dtype)).view(1, -1, kv_heads, attn_head_dim).transpose(1, 2)
                v = F.linear(dropped, p[mp_ + "v_proj.weight"].to(dtype)).view(1, -1, kv_heads, attn_head_dim).transpose(1, 2)
                k = k.repeat_interleave(attn_heads // kv_heads, dim=1)
                v = v.repeat_interleave(attn_heads // kv_heads, dim=1)
                att = ((q @ k.transpose(-1, -2)) / math.sqrt(attn_head_dim)).softmax(-1)          # cross attention, no mask
                merged = F.linear((att @ v).transpose(1, 2).reshape(1, text.shape[1], attn_heads * attn_head_dim),
                                  p[mp_ + "o_proj.weight"].to(dtype))
                text = text + torch.tanh(p["alpha"][st].to(dtype)) * merged                       # :1766-1768
            h = torch.cat([h[:, :vi], h[:, top], text], dim=1)                                    # :1981-1988
            L = h.shape[1]
        x = rmsnorm_ref(h, p[pre + "norm.weight"], eps, dtype)                                    # :941
        sub = {k[len(pre) + 6:]: v for k, v in p.items() if k.startswith(pre + "mixer.")}
        if kind == "M":
            y, _, _ = mixer_forward_ref(sub, x, num_heads=num_heads, head_dim=head_dim, n_groups=n_groups,
                                        ssm_state_size=ssm_state_size, chunk_size=chunk_size, eps=eps,
                                        group_map=group_map, dtype=dtype)
        elif kind == "*":
            q = F.linear(x, sub["q_proj.weight"].to(dtype)).view(b, L, attn_heads, attn_head_dim).transpose(1, 2)
            k = F.linear(x, sub["k_proj.weight"].to(dtype)).view(b, L, kv_heads, attn_head_dim).transpose(1, 2)
            v = F.linear(x, sub["v_proj.weight"].to(dtype)).view(b, L, kv_heads, attn_head_dim).transpose(1, 2)
            k = k.repeat_interleave(attn_heads // kv_heads, dim=1)                                 # repeat_kv :999-1009
            v = v.repeat_interleave(attn_heads // kv_heads, dim=1)
            att = (q @ k.transpose(-1, -2)) / (attn_head_dim ** 0.5)
            att = att.masked_fill(torch.ones(L, L, dtype=torch.bool).triu(1), float("-inf")).softmax(-1)
            y = F.linear((att @ v).transpose(1, 2).reshape(b, L, attn_heads * attn_head_dim), sub["o_proj.weight"].to(dtype))
        elif kind == "-":
            y = F.linear(torch.relu(F.linear(x, sub["up_proj.weight"].to(dtype))) ** 2, sub["down_proj.weight"].to(dtype))
        else:
            raise ValueError(kind)
        h = h + y                                                                                  # :965
    return rmsnorm_ref(h, p["norm_f.weight"], eps, dtype)                                          # :1715


# --------------------------------------------------------------------------------------------
# sequence sharding algebra (new work) -- used to check the multi-GPU path
# --------------------------------------------------------------------------------------------
def fold_boundary_states(local_states: Sequence[torch.Tensor], local_logdecay: Sequence[torch.Tensor],
                         rank: int, initial_states: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Entering state of shard ``rank`` from the per-shard summaries (S_r (b,H,P,N) from zero init,
    log P_r (b,H) = sum of dt*A over the shard):  S_in(r+1) = exp(log P_r) * S_in(r) + S_r."""
    s = (torch.zeros_like(local_states[0]) if initial_states is None else initial_states.clone())
    for r in range(rank):
        s = s * torch.exp(local_logdecay[r])[:, :, None, None].to(s.dtype) + local_states[r]
    return s


def nemotron_random_params(hidden, H, P, G, N, K=4, seed=1234, n_layers=56, nondegenerate=True):
    """Random-init recipe of the reference: ctor modeling_nano.py:414-451 + _init_weights :1339-1383,
    then (nondegenerate) the perturbations of SURVEY.md section 8d config 1 so no term is trivial."""
    g = torch.Generator().manual_seed(seed)
    d_inner, conv_dim = H * P, H * P + 2 * G * N
    W = d_inner + conv_dim + H

    def lin(o, i):
        bound = 1.0 / math.sqrt(i)                      # nn.Linear default: kaiming_uniform(a=sqrt(5))
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound

    p = {"in_proj.weight": lin(W, hidden)}
    kb = 1.0 / math.sqrt(K)                             # depthwise conv: fan_in = K
    p["conv1d.weight"] = (torch.rand(conv_dim, 1, K, generator=g) * 2 - 1) * kb
    p["conv1d.bias"] = (torch.rand(conv_dim, generator=g) * 2 - 1) * kb
    dt = torch.exp(torch.rand(H, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
    p["dt_bias"] = dt + torch.log(-torch.expm1(-dt))
    p["A_log"] = torch.log(torch.arange(1, H + 1, dtype=torch.float32))
    p["D"] = torch.ones(H)
    p["norm.weight"] = torch.ones(d_inner)
    p["out_proj.weight"] = lin(hidden, d_inner) / math.sqrt(n_layers)
    if nondegenerate:
        p["A_log"] = torch.log(torch.rand(H, generator=g) * 15 + 1)
        p["dt_bias"] = torch.randn(H, generator=g) * 0.5 - 2.0
        p["D"] = torch.randn(H, generator=g)
        p["norm.weight"] = 1.0 + 0.1 * torch.randn(d_inner, generator=g)
    return p


def causal_lm_logits_ref(p: dict, input_ids: torch.Tensor, **hybrid_kw):
    """NemotronHForCausalLM.forward (modeling_nano.py:2414-2433): embeddings -> hybrid backbone -> lm_head, fp32 logits of
    every position.  p: the model's state_dict (``backbone.*`` + ``lm_head.weight``)."""
    bb = {k[len("backbone."):]: v for k, v in p.items() if k.startswith("backbone.")}
    dtype = hybrid_kw.get("dtype", torch.float32)
    emb = F.embedding(input_ids, bb["embeddings.weight"].to(dtype))
    h = hybrid_forward_ref(bb, emb, **hybrid_kw)
    return F.linear(h, p["lm_head.weight"].to(dtype)).float()
