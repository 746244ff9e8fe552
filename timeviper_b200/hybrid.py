"""Prefill of the hybrid NemotronH stack: the layer loop of ``NemotronHModel.forward``
(timeviper/model/llm/llm_repo/nano/modeling_nano.py:1550-1746) with the Mamba-2 layers on this package's kernels.

Block = pre-norm + mixer + residual (``NemotronHBlock``, :906-967); mixer by ``hybrid_override_pattern``:
  M  ``Mamba2MixerPrefill``                                   (this package)
  *  ``NemotronHAttention`` (:1012-1117): GQA, no rotary embedding, causal -> library SDPA
  -  ``NemotronHMLP`` (:970-996): down(relu(up(x))^2)          -> cuBLAS
Parameter names are the reference's (``embeddings``, ``layers.N.norm``, ``layers.N.mixer.*``, ``norm_f``), so a reference
``NemotronHModel.state_dict()`` loads strictly.  Only the Mamba-2 mixer is new work; attention and MLP are library calls
kept here so that a whole prefill can be checked against the reference (tests/golden/hybrid_*.npz) and timed.
Deliberately NOT mirrored: the per-layer ``isnan().any()`` host sync (:1690) -- it serialises the GPU every layer."""
import torch
import torch.distributed as dist
from torch import nn

from . import ops as _ops
from .mixer import Mamba2MixerPrefill


class RMSNorm(nn.Module):
    """NemotronHRMSNorm (:888-904): fp32 statistics, fp32 weight multiply, cast back."""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def fusable(self, hidden_states):
        """The one-pass CUDA kernel serves bf16 / fp32 rows of up to 10240 / 5120 elements (``csrc/add_rmsnorm.cu``)."""
        return (hidden_states.is_cuda and hidden_states.dtype in (torch.bfloat16, torch.float32)
                and hidden_states.shape[-1] % (16 // hidden_states.element_size()) == 0
                and hidden_states.shape[-1] <= 128 * 10 * (16 // hidden_states.element_size()))

    def forward(self, hidden_states, residual=None):
        """norm(hidden_states [+ residual]); with a residual returns (normed, hidden_states + residual)."""
        if self.fusable(hidden_states):
            out, summed = _ops.add_rmsnorm(hidden_states, self.weight, self.variance_epsilon, residual)
            return out if residual is None else (out, summed)
        if residual is not None:
            hidden_states = residual + hidden_states
        dtype = hidden_states.dtype
        h = hidden_states.to(torch.float32)
        h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        out = (self.weight.to(torch.float32) * h).to(dtype)
        return out if residual is None else (out, hidden_states)


class Attention(nn.Module):
    def __init__(self, config, layer_idx=None):
        super().__init__()
        self.layer_idx = layer_idx
        self.num_heads, self.num_key_value_heads, self.head_dim = (config.num_attention_heads, config.num_key_value_heads,
                                                                   config.head_dim)
        h = config.hidden_size
        self.q_proj = nn.Linear(h, self.num_heads * self.head_dim, bias=config.attention_bias)
        self.k_proj = nn.Linear(h, self.num_key_value_heads * self.head_dim, bias=config.attention_bias)
        self.v_proj = nn.Linear(h, self.num_key_value_heads * self.head_dim, bias=config.attention_bias)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, h, bias=config.attention_bias)

    def forward(self, hidden_states):
        b, L, _ = hidden_states.shape
        q = self.q_proj(hidden_states).view(b, L, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(hidden_states).view(b, L, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        v = self.v_proj(hidden_states).view(b, L, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        o = nn.functional.scaled_dot_product_attention(q, k, v, is_causal=L > 1,            # :1087-1094
                                                       enable_gqa=self.num_heads != self.num_key_value_heads)
        return self.o_proj(o.transpose(1, 2).reshape(b, L, self.num_heads * self.head_dim))


def cross_attention(attn, hidden_states, encoder_hidden_states):
    """TransV merge module (``Qwen2VLSdpaCrossAttention.forward``, merge_modules/cross_attention.py:225-324, built on the
    NemotronH attention dims): the text tokens (queries) read the vision tokens that a pyramid-drop stage is about to discard
    (keys / values); no mask, no rotary embedding, library SDPA."""
    b, Lq, _ = hidden_states.shape
    Lk = encoder_hidden_states.shape[1]
    nh, nkv, d = attn.num_heads, attn.num_key_value_heads, attn.head_dim
    q = attn.q_proj(hidden_states).view(b, Lq, nh, d).transpose(1, 2)
    k = attn.k_proj(encoder_hidden_states).view(b, Lk, nkv, d).transpose(1, 2)
    v = attn.v_proj(encoder_hidden_states).view(b, Lk, nkv, d).transpose(1, 2)
    o = nn.functional.scaled_dot_product_attention(q, k, v, is_causal=False, enable_gqa=nh != nkv)
    return attn.o_proj(o.transpose(1, 2).reshape(b, Lq, nh * d))


_two_part_ok = True


def _two_part_attention(q, k, v, L, rep):
    """Lower-right causal attention of this rank's L queries over (r+1)*L keys as TWO library calls on their fast paths --
    full attention over the r*L keys of the earlier ranks and square causal attention over this rank's own keys -- merged
    through the log-sum-exp each call returns.  (A lower-right CausalBias makes PyTorch pick the FlashAttention-2 backend,
    which on B200 is ~3x slower than the cuDNN one it uses for plain causal attention: 130 ms against 42 ms of work at
    40K tokens per rank.)  Uses the private aten cuDNN SDPA op for the LSE output; returns None if it is not available, and
    the caller falls back to the single masked call."""
    global _two_part_ok
    if not _two_part_ok:
        return None
    try:
        op = torch.ops.aten._scaled_dot_product_cudnn_attention
        kp, vp = k[:, :, :-L], v[:, :, :-L]
        ko, vo = k[:, :, -L:], v[:, :, -L:]
        if rep > 1:       # the private op has no enable_gqa: expand the KV heads (views, no copy)
            ex = lambda t: t.unsqueeze(2).expand(-1, -1, rep, -1, -1).reshape(t.shape[0], t.shape[1] * rep, t.shape[2], t.shape[3])
            kp, vp, ko, vo = ex(kp), ex(vp), ex(ko), ex(vo)
        oa, la = op(q, kp, vp, None, True, 0.0, False)[:2]
        ob, lb = op(q, ko, vo, None, True, 0.0, True)[:2]
        la, lb = la.float().reshape(la.shape[0], la.shape[1], -1, 1), lb.float().reshape(lb.shape[0], lb.shape[1], -1, 1)
        m = torch.maximum(la, lb)
        wa, wb = torch.exp(la - m), torch.exp(lb - m)
        return ((oa.float() * wa + ob.float() * wb) / (wa + wb)).to(q.dtype)
    except Exception:       # op missing / shape unsupported on this build: one-time fallback
        _two_part_ok = False
        return None


def shard_bounds(total, world):
    """Balanced contiguous split of `total` tokens over `world` ranks: offsets [0, ..., total] (first ranks one longer)."""
    base, rem = divmod(total, world)
    offs = [0]
    for r in range(world):
        offs.append(offs[-1] + base + (1 if r < rem else 0))
    return offs


def _gather_lengths(L, device, group):
    """Shard lengths of every rank (host ints; one small all-gather + sync, once per forward or per drop)."""
    world = dist.get_world_size(group)
    t = torch.tensor([L], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return [int(v) for v in out.tolist()]


def sharded_attention_forward(attn, hidden_states, group=None, lens=None):
    """Causal attention of a sequence-sharded layer: rank r holds a contiguous shard (lengths `lens`, default: all equal to
    this rank's).  K and V of every rank are all-gathered (GQA: 2 x kv_heads x head_dim values per token, 4 KB at the 9B
    shape -- 1/10 of the hidden state), each rank attends its queries over the keys of ranks <= r with a LOWER-RIGHT
    aligned causal mask (its queries are the last positions of that key range).  The work per rank grows with r (causal);
    a zig-zag split would balance it but breaks the contiguous shards the Mamba-2 layers need."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, L, _ = hidden_states.shape
    if lens is None:
        lens = [L] * world
    Lmax = max(lens)
    nh, nkv, d = attn.num_heads, attn.num_key_value_heads, attn.head_dim
    q = attn.q_proj(hidden_states).view(b, L, nh, d).transpose(1, 2)
    kv = torch.stack([attn.k_proj(hidden_states), attn.v_proj(hidden_states)])                       # (2, b, L, nkv*d)
    if L < Lmax:
        kv = nn.functional.pad(kv, (0, 0, 0, Lmax - L))
    kv = kv.contiguous()
    gathered = torch.empty((world,) + tuple(kv.shape), dtype=kv.dtype, device=kv.device)
    dist.all_gather_into_tensor(gathered.view(-1), kv.view(-1), group=group)
    parts = [gathered[j, :, :, :lens[j]] for j in range(rank + 1)]                                    # keys this rank may see
    kv_all = parts[0] if rank == 0 else torch.cat(parts, dim=2)                                       # (2, b, Lk, nkv*d)
    Lk = kv_all.shape[2]
    k = kv_all[0].view(b, Lk, nkv, d).transpose(1, 2)
    v = kv_all[1].view(b, Lk, nkv, d).transpose(1, 2)
    o = None
    if hidden_states.is_cuda and rank > 0 and q.dtype in (torch.bfloat16, torch.float16):   # the cuDNN op is half-precision only
        o = _two_part_attention(q, k, v, L, nh // nkv)
    if o is None:
        if hidden_states.is_cuda:
            from torch.nn.attention.bias import causal_lower_right
            mask = causal_lower_right(L, Lk)
        else:       # CPU (gloo tests of the host logic): the same mask, materialised
            mask = torch.ones(L, Lk, dtype=torch.bool).tril(diagonal=Lk - L)
        o = nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, enable_gqa=nh != nkv)
    return attn.o_proj(o.transpose(1, 2).reshape(b, L, nh * d))


class MLP(nn.Module):
    def __init__(self, config, layer_idx=None):
        super().__init__()
        if config.mlp_hidden_act != "relu2":
            raise NotImplementedError("NemotronH uses the squared-ReLU MLP")
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size_mlp, bias=config.mlp_bias)
        self.down_proj = nn.Linear(config.intermediate_size_mlp, config.hidden_size, bias=config.mlp_bias)

    def forward(self, x):
        return self.down_proj(torch.square(torch.relu(self.up_proj(x))))


_KIND = {"M": "mamba", "*": "attention", "-": "mlp"}


class HybridBlock(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        self.block_type = _KIND[config.hybrid_override_pattern[layer_idx]]
        self.residual_in_fp32 = config.residual_in_fp32
        self.norm = RMSNorm(config.hidden_size, eps=config.layer_norm_epsilon)
        self.mixer = {"mamba": Mamba2MixerPrefill, "attention": Attention, "mlp": MLP}[self.block_type](config, layer_idx)

    def forward(self, hidden_states, cache_params=None, cache_position=None, group=None, mixer_ops=None, lens=None,
                delta=None, defer_add=False):
        """group: the sequence is sharded contiguously over this process group (rank order = token order; `lens`: the shard
        lengths when they are not all equal)."""
        if delta is not None:       # the previous block's output not yet added to its residual: add + norm in one pass
            h, hidden_states = self.norm(delta, residual=hidden_states)
            residual = hidden_states
        else:
            residual = hidden_states.to(torch.float32) if self.residual_in_fp32 else hidden_states
            h = self.norm(hidden_states.to(self.norm.weight.dtype))
        sharded = group is not None and dist.get_world_size(group) > 1
        if self.block_type == "mamba":
            if sharded:
                from .sharded import sharded_mixer_forward
                kw = {} if mixer_ops is None else {"ops": mixer_ops}
                h = sharded_mixer_forward(self.mixer, h, group=group, cache_params=cache_params, **kw)
            else:
                h = self.mixer(h, cache_params=cache_params, cache_position=cache_position)
        elif self.block_type == "attention" and sharded:
            h = sharded_attention_forward(self.mixer, h, group, lens)
        else:
            h = self.mixer(h)
        if defer_add:               # the caller fuses `residual + h` into the next norm (HybridPrefillStack.forward)
            return residual, h
        return residual + h


def parse_pdrop_type(pdrop_type):
    """'type_layer_ratio-...' -> (types, layers, ratios with a leading 1)   (modeling_nano.py:1469-1477; the default
    schedule of evaluate.py:167-172 is 'uni_14_0.8-attn_21_0.6-attn_30_0.4-attn_39_0.2')."""
    parts = [t.split("_") for t in pdrop_type.split("-")]
    if not all(len(t) == 3 for t in parts):
        raise ValueError("pdrop_type should be like 'type_layernum_ratio-...'")
    return [t[0] for t in parts], [int(t[1]) for t in parts], [1.0] + [float(t[2]) for t in parts]


@torch.no_grad()
def pdrop_select(h, stage, kind, ratios, attn, vision_index, num_vision_tokens, text_prompt_len):
    """TransV / pyramid-drop token selection for inference, one sample (``pdrop_no_pack``, modeling_nano.py:1779-1988):
    sorted sequence indices of the vision tokens that survive stage ``stage`` and the index of the first token after
    the vision block.  h: (L, hidden), the hidden states entering the layer (NOT normed, as in the reference :1821-1836).

    'uni': evenly spaced (:1950-1957).  'attn': attention of the LAST prompt token over the sequence, with this attention
    layer's q/k projections, softmax in fp32, mean over heads, top-k among the vision tokens (:1917-1947).  The reference
    projects q for every position and builds an (L, L) mask; only one query row is ever read, so this computes that row:
    one (1 x hidden) q GEMV, the k projection of the positions up to the query, and heads x L scores."""
    image_tokens = int(num_vision_tokens * ratios[stage])
    keep = int(num_vision_tokens * ratios[stage + 1])
    if "attn" in kind:
        if attn is None:
            raise ValueError("an attention-ranked drop must sit on an attention layer (modeling_nano.py:1824)")
        pq = text_prompt_len + image_tokens - 1
        nh, nkv, d = attn.num_heads, attn.num_key_value_heads, attn.head_dim
        q = attn.q_proj(h[pq:pq + 1]).view(nh, 1, d)
        k = attn.k_proj(h[:pq + 1]).view(pq + 1, nkv, d).transpose(0, 1)                  # keys the causal row can see
        k = k.repeat_interleave(nh // nkv, dim=0)
        w = torch.softmax((q @ k.transpose(1, 2)) / (d ** 0.5), dim=-1, dtype=torch.float32).to(h.dtype)   # (nh, 1, pq+1)
        w = w.mean(0)[0, vision_index:vision_index + image_tokens]
        top = w.topk(keep).indices
    elif "uni" in kind:
        top = torch.linspace(0, image_tokens - 1, keep, dtype=torch.long).to(h.device)     # on the host, as the reference's
    else:
        raise NotImplementedError(kind)
    return (top + vision_index).sort().values, vision_index + image_tokens


@torch.no_grad()
def sharded_pdrop(h, offs, group, stage, kind, ratios, attn, vision_index, num_vision_tokens, text_prompt_len, trace=None):
    """``pdrop_select`` + the token drop for ONE sample whose sequence is sharded over `group` (this rank holds the global
    positions offs[rank] .. offs[rank+1]).  Returns this rank's shard of the shortened sequence, re-balanced
    (``shard_bounds``), and the new offsets.  Same rule as the unsharded path: 'uni' needs no communication; 'attn' broadcasts
    the query row of the last prompt token, every rank scores its own keys, the softmax is normalised with the global
    maximum / sum (two small all-reduces), and the per-token weights of the vision block (4 bytes per token) are gathered so
    that every rank takes the same top-k.  The surviving tokens then move with ONE all-to-all."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev, dtype = h.device, h.dtype
    image_tokens = int(num_vision_tokens * ratios[stage])
    keep = int(num_vision_tokens * ratios[stage + 1])
    lo, hi = offs[rank], offs[rank + 1]
    total = offs[-1]
    if "attn" in kind:
        if attn is None:
            raise ValueError("an attention-ranked drop must sit on an attention layer (modeling_nano.py:1824)")
        pq = text_prompt_len + image_tokens - 1                                # global position of the last prompt token
        owner = max(r for r in range(world) if offs[r] <= pq)
        nh, nkv, d = attn.num_heads, attn.num_key_value_heads, attn.head_dim
        q = attn.q_proj(h[0, pq - lo:pq - lo + 1]).view(nh, 1, d) if rank == owner else torch.empty(nh, 1, d, dtype=dtype, device=dev)
        dist.broadcast(q, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        n_loc = max(0, min(hi, pq + 1) - lo)                                   # my keys the causal row can see
        k = attn.k_proj(h[0, :n_loc]).view(n_loc, nkv, d).transpose(0, 1).repeat_interleave(nh // nkv, dim=0)
        sc = ((q @ k.transpose(1, 2)) / (d ** 0.5))[:, 0].float()              # (nh, n_loc), scores rounded to dtype first
        m = sc.max(dim=-1).values if n_loc > 0 else torch.full((nh,), float("-inf"), device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        e = torch.exp(sc - m[:, None])
        z = e.sum(-1)
        dist.all_reduce(z, op=dist.ReduceOp.SUM, group=group)
        w = (e / z[:, None]).to(dtype).mean(0)                                 # (n_loc,): mean over heads, in dtype
        # my part of the vision block [vision_index, vision_index + image_tokens)
        a, b_ = max(lo, vision_index), min(hi, vision_index + image_tokens, pq + 1)
        mine = w[a - lo:b_ - lo] if b_ > a else w[:0]
        cnt = [max(0, min(offs[r + 1], vision_index + image_tokens) - max(offs[r], vision_index)) for r in range(world)]
        pad = torch.zeros(max(cnt), dtype=dtype, device=dev)
        pad[:mine.numel()] = mine
        allw = torch.empty(world, max(cnt), dtype=dtype, device=dev)
        dist.all_gather_into_tensor(allw.view(-1), pad, group=group)
        wfull = torch.cat([allw[r, :cnt[r]] for r in range(world)])            # (image_tokens,), identical on every rank
        top = wfull.topk(keep).indices.cpu()
    elif "uni" in kind:
        top = torch.linspace(0, image_tokens - 1, keep, dtype=torch.long)
    else:
        raise NotImplementedError(kind)
    start = vision_index + image_tokens
    kept = torch.cat([torch.arange(0, vision_index), (top + vision_index).sort().values, torch.arange(start, total)])   # host
    if trace is not None:
        trace.append((top + vision_index).sort().values.clone())
    new_offs = shard_bounds(kept.numel(), world)
    # kept is sorted: the tokens of old shard s are the new positions [pos[s], pos[s+1])
    pos = torch.searchsorted(kept, torch.tensor(offs)).tolist()
    a, b_ = pos[rank], pos[rank + 1]
    send = h[:, (kept[a:b_] - lo).to(dev)].contiguous()[0]                     # (b_-a, hidden)
    ovl = lambda x0, x1, y0, y1: max(0, min(x1, y1) - max(x0, y0))
    in_splits = [ovl(a, b_, new_offs[r], new_offs[r + 1]) for r in range(world)]
    out_splits = [ovl(pos[s_], pos[s_ + 1], new_offs[rank], new_offs[rank + 1]) for s_ in range(world)]
    out = torch.empty(sum(out_splits), h.shape[-1], dtype=dtype, device=dev)
    dist.all_to_all_single(out, send, out_splits, in_splits, group=group)
    return out.unsqueeze(0), new_offs


class HybridPrefillStack(nn.Module):
    def __init__(self, config):
        super().__init__()
        if len(config.hybrid_override_pattern) != config.num_hidden_layers:
            raise ValueError("hybrid_override_pattern needs one character per layer")
        self.config = config
        self.embeddings = nn.Embedding(config.vocab_size, config.hidden_size)
        self.layers = nn.ModuleList([HybridBlock(config, i) for i in range(config.num_hidden_layers)])
        self.norm_f = RMSNorm(config.hidden_size, eps=config.layer_norm_epsilon)
        # TransV merge modules (modeling_nano.py:1481-1523): one cross-attention per pyramid-drop stage + the gates alpha
        self.merge_modules, self.alpha = None, None
        if getattr(config, "merge_module", "no_merge") == "CrossAttention":
            if not getattr(config, "pdrop_type", None):
                raise ValueError("merge_module='CrossAttention' needs config.pdrop_type (one module per drop stage)")
            kinds, drop_layers, _ = parse_pdrop_type(config.pdrop_type)
            if any("drop" in k for k in kinds):
                raise NotImplementedError("stages of the '...drop' kind (no merge at that stage) are not wired")
            self.merge_modules = nn.ModuleList([Attention(config, layer_idx=i) for i in drop_layers])
            self.alpha = nn.Parameter(torch.zeros(len(drop_layers)))
        elif getattr(config, "merge_module", "no_merge") != "no_merge":
            raise ValueError(f"Invalid merge module name: {config.merge_module}")

    @torch.no_grad()
    def forward(self, input_ids=None, inputs_embeds=None, cache_params=None, pdrop=None, group=None, mixer_ops=None):
        """Prefill: (b, L) token ids or (b, L, hidden) embeddings -> last hidden states (b, L', hidden) after ``norm_f``.
        ``pdrop`` (the reference's ``train_pdrop_args`` + ``config.pdrop_type``) = dict(pdrop_type=
        'uni_14_0.8-attn_21_0.6-...', first_vision_token_position, num_vision_tokens, text_prompt_len): TransV /
        pyramid-drop of vision tokens before the listed layers (modeling_nano.py:1634-1666; batch 1, ``no_merge``), so the
        layers after a stage see L' < L tokens: [tokens before the video | surviving vision tokens | text].
        ``group`` (a torch.distributed process group of more than one rank): the inputs are this rank's contiguous shard of
        ONE sequence (equal shards, rank order = token order) and so is the result; Mamba-2 layers run
        ``sharded_mixer_forward`` (conv halo + boundary-state exchange), attention layers all-gather K / V
        (``sharded_attention_forward``), everything else is token-local.  With ``pdrop`` the positions in it are GLOBAL
        (whole-sequence) positions; after every drop the survivors are re-balanced over the ranks (``sharded_pdrop``).
        ``cache_params`` (reference cache interface) receives the conv / SSM states of every Mamba-2 layer.
        Limits (stated, not hidden): no ``attention_mask`` (batch 1 or unpadded batches only) and the attention layers do
        not write a KV cache, so this stack prefills and scores the last position; token-by-token decode after it needs
        the reference's attention cache and is not wired here."""
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("exactly one of input_ids / inputs_embeds")
        h = self.embeddings(input_ids) if inputs_embeds is None else inputs_embeds
        pos = torch.arange(h.shape[1])          # on the HOST: the mixer branches on cache_position[0] > 0 (no device sync)
        sharded = group is not None and dist.get_world_size(group) > 1
        lens = offs = None
        if pdrop is not None:
            if h.shape[0] != 1:
                raise NotImplementedError("pyramid-drop is wired for batch 1 (the reference's inference path)")
            kinds, drop_layers, ratios = parse_pdrop_type(pdrop["pdrop_type"])
        if sharded and pdrop is not None:
            lens = _gather_lengths(h.shape[1], h.device, group)
            offs = [0]
            for n_ in lens:
                offs.append(offs[-1] + n_)
        # `residual + mixer output` of block i is folded into the pre-norm of block i+1 (one pass instead of ~10 elementwise
        # ones) whenever the fused kernel serves the dtype; a drop stage needs the materialised sum, so it adds first
        fuse = (not self.config.residual_in_fp32) and self.norm_f.fusable(h)
        delta = None
        for i, layer in enumerate(self.layers):
            if pdrop is not None and i in drop_layers:
                if delta is not None:
                    h, delta = h + delta, None
                st = drop_layers.index(i)
                vi = pdrop["first_vision_token_position"]
                att = layer.mixer if layer.block_type == "attention" else None
                if sharded:
                    if self.merge_modules is not None:
                        raise NotImplementedError("the TransV merge module over a sequence-sharded sample is not built")
                    h, offs = sharded_pdrop(h, offs, group, st, kinds[st], ratios, att, vi, pdrop["num_vision_tokens"],
                                            pdrop["text_prompt_len"], trace=pdrop.get("_trace"))
                    lens = [offs[r + 1] - offs[r] for r in range(len(offs) - 1)]
                else:
                    top, start = pdrop_select(h[0], st, kinds[st], ratios, att, vi, pdrop["num_vision_tokens"],
                                              pdrop["text_prompt_len"])
                    text = h[:, start:]
                    if self.merge_modules is not None:                                     # TransV: merge_dropped_information
                        keep_mask = torch.ones(start - vi, dtype=torch.bool, device=h.device)
                        keep_mask[top - vi] = False
                        text = text + torch.tanh(self.alpha[st]).to(h.dtype) * cross_attention(
                            self.merge_modules[st], text, h[:, vi:start][:, keep_mask])
                    h = torch.cat([h[:, :vi], h[:, top], text], dim=1)                     # :1981-1988
                    if pdrop.get("_trace") is not None:     # tests: the surviving vision positions of this stage
                        pdrop["_trace"].append(top.cpu().clone())
                pos = torch.arange(h.shape[1])
            r = layer(h, cache_params=cache_params, cache_position=pos, group=group if sharded else None,
                      mixer_ops=mixer_ops, lens=lens, delta=delta, defer_add=fuse)
            h, delta = r if fuse else (r, None)
        if delta is not None:
            return self.norm_f(delta, residual=h)[0]
        return self.norm_f(h)


class HybridCausalLM(nn.Module):
    """``NemotronHForCausalLM`` for prefill (modeling_nano.py:2286-2292, forward :2414-2433): ``backbone`` + ``lm_head`` with
    the reference's parameter names, so its ``state_dict`` loads strictly.

    The reference projects EVERY position to the vocabulary and upcasts to fp32 (:2433) -- at 128K video tokens and the
    131,072-entry Nanov2 vocabulary that is a 64 GiB tensor of which generation reads one row.  ``forward`` therefore returns
    the fp32 logits of the LAST position by default (``(b, 1, vocab)``: the reference's ``logits[:, -1:]`` up to the
    summation order of the GEMM); ``all_positions=True`` gives the reference's full ``(b, L, vocab)`` tensor."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.backbone = HybridPrefillStack(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)

    @torch.no_grad()
    def forward(self, input_ids=None, inputs_embeds=None, cache_params=None, all_positions=False, pdrop=None, group=None):
        """With ``group`` (sequence-sharded prefill): every rank returns the logits of the sequence's LAST position (computed
        on the last rank and broadcast); ``all_positions`` then gives the logits of this rank's shard."""
        h = self.backbone(input_ids=input_ids, inputs_embeds=inputs_embeds, cache_params=cache_params, pdrop=pdrop, group=group)
        if not all_positions:
            h = h[:, -1:]
        logits = self.lm_head(h.to(self.lm_head.weight.dtype)).float()
        if group is not None and dist.get_world_size(group) > 1 and not all_positions:
            dist.broadcast(logits, src=dist.get_global_rank(group, dist.get_world_size(group) - 1), group=group)
        return logits
