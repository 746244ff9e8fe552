"""Reference-facing operators of the Mamba-2 mixer prefill path, backed by hand-written sm_100a CUDA.

Same names, argument meaning and error behaviour as the third-party functions the reference binds at
timeviper/model/llm/llm_repo/nano/modeling_nano.py:60-82:

  causal_conv1d_fn           causal_conv1d 1.5.2   (call site modeling_nano.py:619-624)
  mamba_chunk_scan_combined  mamba_ssm 2.2.5       (signature visualize/nano/my_ssd_combined.py:1270-1306,
                                                    call site modeling_nano.py:639-653)
  rmsnorm_fn                 mamba_ssm layernorm_gated (call site modeling_nano.py:372-380)

PyTorch is plumbing here: it owns device memory and the stream; every arithmetic step runs in
libtimeviper_b200.so through the C ABI of include/timeviper_b200.h.  No CPU path exists.
"""
import ctypes as C
import math

import torch

from . import _lib as L

_byref = C.byref       # `C` is also the name of the SSM output-projection argument in the reference signatures

_DT = {torch.float32: L.TV_F32, torch.bfloat16: L.TV_BF16}


def _dtype_code(t, what):
    if t.dtype not in _DT:
        raise ValueError(f"{what}: dtype {t.dtype} unsupported (float32, bfloat16)")
    return _DT[t.dtype]


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("timeviper_b200 ops run on CUDA tensors only (no CPU fallback)")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# ------------------------------------------------------------------------------------------------
def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                     final_states_out=None, activation=None):
    """x: (batch, dim, seqlen) -- channel-last (x.stride(1) == 1), possibly a strided view;
    weight: (dim, width); bias: (dim,); initial_states: (batch, dim, width-1);
    activation: None | "silu" | "swish".  Returns out (batch, dim, seqlen) channel-last
    [, final_states (batch, dim, width-1)]."""
    return causal_conv1d_into(None, x, weight, bias, seq_idx, initial_states, return_final_states,
                              final_states_out, activation)


def causal_conv1d_into(_out, x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                       final_states_out=None, activation=None):
    """causal_conv1d_fn writing into `_out`, a caller-owned (batch, seqlen, dim) buffer (None: allocate)."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    if seq_idx is not None:
        raise NotImplementedError("causal_conv1d_fn: seq_idx (packed sequences) is outside the prefill path")
    if x.dim() != 3:
        raise ValueError("causal_conv1d_fn: x must be (batch, dim, seqlen)")
    _require_cuda(x, weight, bias, initial_states)
    b, dim, seqlen = x.shape
    if weight.shape[0] != dim or weight.dim() != 2:
        raise ValueError(f"causal_conv1d_fn: weight {tuple(weight.shape)} does not match dim {dim}")
    width = weight.shape[1]
    vec = 16 // x.element_size()               # the kernel moves 16-byte pieces of a channel-last row
    if x.stride(1) != 1 or x.stride(2) % vec or x.stride(0) % vec or x.data_ptr() % 16:
        # upstream also accepts channel-first; a view whose rows are not 16-byte aligned (e.g. a split of an in_proj
        # output of odd width) is repacked as well
        x = x.transpose(1, 2).contiguous().transpose(1, 2)
    weight = weight.to(x.dtype).contiguous()
    bias = None if bias is None else bias.to(x.dtype).contiguous()
    if initial_states is not None:
        if tuple(initial_states.shape) != (b, dim, width - 1):
            raise ValueError("causal_conv1d_fn: initial_states must be (batch, dim, width-1)")
        initial_states = initial_states.to(x.dtype).contiguous()
    if _out is not None:
        if tuple(_out.shape) != (b, seqlen, dim) or _out.dtype != x.dtype or _out.stride(2) != 1:
            raise ValueError("causal_conv1d_fn: _out must be (batch, seqlen, dim) of x's dtype, unit channel stride")
        out = _out
    else:
        out = torch.empty((b, seqlen, dim), dtype=x.dtype, device=x.device)
    fin = None
    if return_final_states or final_states_out is not None:
        if final_states_out is not None:
            if (tuple(final_states_out.shape) != (b, dim, width - 1) or not final_states_out.is_contiguous()
                    or final_states_out.dtype != x.dtype):
                raise ValueError("causal_conv1d_fn: final_states_out must be contiguous (batch, dim, width-1)")
            fin = final_states_out
        else:
            fin = torch.empty((b, dim, width - 1), dtype=x.dtype, device=x.device)
    p = L.ConvParams(x=_ptr(x), weight=_ptr(weight), bias=_ptr(bias), initial_states=_ptr(initial_states),
                     out=_ptr(out), final_states=_ptr(fin), batch=b, dim=dim, seqlen=seqlen, width=width,
                     x_batch_stride=x.stride(0), x_seq_stride=x.stride(2),
                     out_batch_stride=out.stride(0), out_seq_stride=out.stride(1),
                     silu=int(activation in ("silu", "swish")), dtype=_dtype_code(x, "causal_conv1d_fn"))
    L.check(L.load().tv_causal_conv1d_fwd(C.byref(p), _stream(x)), "causal_conv1d_fn")
    out = out.transpose(1, 2)
    return (out, fin) if return_final_states else out


# ------------------------------------------------------------------------------------------------
def rmsnorm_fn(x, weight, bias, z=None, eps=1e-6, group_size=None, norm_before_gate=True, upcast=True):
    """Gated grouped RMSNorm: x, z (..., d); weight/bias (d,).  Math is always fp32 (`upcast`)."""
    _require_cuda(x, weight, bias, z)
    shape = x.shape
    d = shape[-1]
    vec = 16 // x.element_size()               # the kernel moves 16-byte pieces of a row

    def aligned(t):                            # unit stride along d, 16-byte aligned rows; otherwise repack
        return t if (t.stride(-1) == 1 and t.stride(0) % vec == 0 and t.data_ptr() % 16 == 0) else t.contiguous()

    x2 = aligned(x.reshape(-1, d))
    z2 = None
    if z is not None:
        if z.shape != shape:
            raise ValueError("rmsnorm_fn: z must have the shape of x")
        z2 = aligned(z.to(x.dtype).reshape(-1, d))
    if weight.shape != (d,):
        raise ValueError("rmsnorm_fn: weight must be (d,)")
    weight = weight.to(x.dtype).contiguous()
    bias = None if bias is None else bias.to(x.dtype).contiguous()
    g = d if group_size is None else int(group_size)
    out = torch.empty((x2.shape[0], d), dtype=x.dtype, device=x.device)
    if x2.shape[0] == 0:
        return out.reshape(shape)
    p = L.RmsnormParams(x=_ptr(x2), z=_ptr(z2), weight=_ptr(weight), bias=_ptr(bias), out=_ptr(out),
                        rows=x2.shape[0], d=d, group_size=g, x_row_stride=x2.stride(0),
                        z_row_stride=0 if z2 is None else z2.stride(0), out_row_stride=out.stride(0),
                        eps=float(eps), norm_before_gate=int(bool(norm_before_gate)),
                        dtype=_dtype_code(x, "rmsnorm_fn"))
    L.check(L.load().tv_gated_rmsnorm_fwd(C.byref(p), _stream(x)), "rmsnorm_fn")
    return out.reshape(shape)


# ------------------------------------------------------------------------------------------------
_workspaces = {}


def _workspace(nbytes, device, owner_stream=None):
    """Grow-only scratch per (device, stream): the .so never allocates (ownership).
    `owner_stream`: the stream whose scratch to use when a call is enqueued on a helper stream."""
    owner = owner_stream if owner_stream is not None else torch.cuda.current_stream(device)
    key = (device.index, owner.cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


force_simt_default = False  # tests: route every SSD call to the fp32 CUDA-core family (an fp32-accumulating reference on the GPU)
simt_fallbacks = 0          # bf16 SSD calls served by the fp32 CUDA-core family (see _warn_simt_fallback)
_simt_warned = set()


def _warn_simt_fallback(headdim, dstate, chunk_size):
    """The tcgen05 kernel is specialised for (bf16, headdim 80, dstate 128, chunk 128) with 16-byte aligned rows; every other
    bf16 call runs on the CUDA-core family, which is 50-60x slower at the 9B geometry.  Say so once per shape and count."""
    global simt_fallbacks
    simt_fallbacks += 1
    key = (headdim, dstate, chunk_size)
    if key not in _simt_warned:
        _simt_warned.add(key)
        import warnings
        warnings.warn(f"timeviper_b200: bf16 SSD scan with headdim={headdim}, dstate={dstate}, chunk_size={chunk_size} "
                      "(or unaligned views) is not served by the tcgen05 kernel (80/128/128); using the fp32 CUDA-core "
                      "kernels, which are far slower at large sizes", RuntimeWarning, stacklevel=4)


def launch_count():
    """Kernels libtimeviper_b200.so has enqueued so far in this process (all ops, all streams)."""
    return int(L.load().tv_debug_launch_count())


def _ssd_call(x, dt, A, B, C_, chunk_size, D, z, dt_bias, initial_states, dt_softplus, dt_limit, mode,
              force_simt=False, want_final=True, want_logdecay=False, reuse_dt_cumsum=False, final_out=None,
              logdecay_out=None, workspace_stream=None):
    batch, seqlen, nheads, headdim = x.shape
    ngroups, dstate = B.shape[2], B.shape[3]
    assert nheads % ngroups == 0
    assert B.shape == (batch, seqlen, ngroups, dstate)
    assert dt.shape == (batch, seqlen, nheads)
    assert A.shape == (nheads,)
    if C_ is not None:
        assert C_.shape == B.shape
    if z is not None:
        assert z.shape == x.shape
    if D is not None:
        assert D.shape == (nheads, headdim) or D.shape == (nheads,)
    if initial_states is not None:
        assert initial_states.shape == (batch, nheads, headdim, dstate)
    _require_cuda(x, dt, A, B, C_, D, z, dt_bias, initial_states)
    code = _dtype_code(x, "mamba_chunk_scan_combined")
    dev = x.device
    # my_ssd_combined.py:775-788: make the last dim contiguous where the kernels need it
    if x.stride(-1) != 1:
        x = x.contiguous()
    B = B.to(x.dtype)
    if B.stride(-1) != 1:
        B = B.contiguous()
    if C_ is not None:
        C_ = C_.to(x.dtype)
        if C_.stride(-1) != 1:
            C_ = C_.contiguous()
    if z is not None:
        z = z.to(x.dtype)
        if z.stride(-1) != 1:
            z = z.contiguous()
    dt = dt.to(x.dtype)
    A32 = A.detach().to(torch.float32).contiguous()
    D32 = None if D is None else D.detach().to(torch.float32).contiguous()
    bias32 = None if dt_bias is None else dt_bias.detach().to(torch.float32).contiguous()
    init32 = None if initial_states is None else initial_states.detach().to(torch.float32).contiguous()
    lo, hi = float(dt_limit[0]), float(dt_limit[1])
    out = None
    if mode == L.TV_SSD_FULL:
        out = torch.empty((batch, seqlen, nheads, headdim), dtype=x.dtype, device=dev)
    fin = logdecay = None
    if want_final:
        fin = final_out if final_out is not None else torch.empty((batch, nheads, headdim, dstate),
                                                                  dtype=torch.float32, device=dev)
        assert fin.shape == (batch, nheads, headdim, dstate) and fin.dtype == torch.float32 and fin.is_contiguous()
    if want_logdecay:
        logdecay = logdecay_out if logdecay_out is not None else torch.empty((batch, nheads), dtype=torch.float32,
                                                                             device=dev)
        assert logdecay.shape == (batch, nheads) and logdecay.dtype == torch.float32 and logdecay.is_contiguous()
    p = L.SsdParams(
        x=_ptr(x), dt=_ptr(dt), A=_ptr(A32), B=_ptr(B), C=_ptr(C_), D=_ptr(D32), z=_ptr(z),
        dt_bias=_ptr(bias32), initial_states=_ptr(init32), out=_ptr(out), final_states=_ptr(fin),
        logdecay_sum=_ptr(logdecay), batch=batch, seqlen=seqlen, nheads=nheads, headdim=headdim,
        ngroups=ngroups, dstate=dstate, chunk_size=int(chunk_size),
        x_batch_stride=x.stride(0), x_seq_stride=x.stride(1), x_head_stride=x.stride(2),
        dt_batch_stride=dt.stride(0), dt_seq_stride=dt.stride(1), dt_head_stride=dt.stride(2),
        b_batch_stride=B.stride(0), b_seq_stride=B.stride(1), b_group_stride=B.stride(2),
        c_batch_stride=0 if C_ is None else C_.stride(0), c_seq_stride=0 if C_ is None else C_.stride(1),
        c_group_stride=0 if C_ is None else C_.stride(2),
        z_batch_stride=0 if z is None else z.stride(0), z_seq_stride=0 if z is None else z.stride(1),
        z_head_stride=0 if z is None else z.stride(2),
        d_has_hdim=int(D is not None and D.dim() == 2), dt_softplus=int(bool(dt_softplus)),
        dt_min=lo, dt_max=hi if math.isfinite(hi) else float("inf"), dtype=code, mode=mode,
        force_simt=int(bool(force_simt or force_simt_default)), reuse_dt_cumsum=int(bool(reuse_dt_cumsum)))
    lib = L.load()
    if code == L.TV_BF16 and not (force_simt or force_simt_default) and lib.tv_ssd_kernel_family(C.byref(p)) == 0:
        _warn_simt_fallback(headdim, dstate, int(chunk_size))
    need = lib.tv_ssd_workspace_bytes(C.byref(p))
    ws = _workspace(need, dev, workspace_stream)
    L.check(lib.tv_ssd_chunk_scan_fwd(C.byref(p), _ptr(ws), ws.numel(), _stream(x)), "mamba_chunk_scan_combined")
    return out, fin, logdecay


def mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None,
                              seq_idx=None, cu_seqlens=None, dt_softplus=False, dt_limit=(0.0, float("inf")),
                              return_final_states=False, return_varlen_states=False, _force_simt=False,
                              _reuse_dt_cumsum=False):
    """
    Argument:
        x: (batch, seqlen, nheads, headdim)        dt: (batch, seqlen, nheads)       A: (nheads)
        B, C: (batch, seqlen, ngroups, dstate)     D: (nheads, headdim) or (nheads,) z: (batch, seqlen, nheads, headdim)
        dt_bias: (nheads,)                         initial_states: (batch, nheads, headdim, dstate)
    Return:
        out: (batch, seqlen, nheads, headdim) [, final_states (batch, nheads, headdim, dstate) fp32]
    """
    if seq_idx is not None or cu_seqlens is not None or return_varlen_states:
        raise NotImplementedError("mamba_chunk_scan_combined: seq_idx / cu_seqlens / varlen states are "
                                  "outside the prefill path (the reference passes seq_idx=None)")
    out, fin, _ = _ssd_call(x, dt, A, B, C, chunk_size, D, z, dt_bias, initial_states, dt_softplus, dt_limit,
                            L.TV_SSD_FULL, force_simt=_force_simt, want_final=return_final_states,
                            reuse_dt_cumsum=_reuse_dt_cumsum)
    return (out, fin) if return_final_states else out


def mamba_chunk_state_summary(x, dt, A, B, chunk_size, dt_bias=None, dt_softplus=False,
                              dt_limit=(0.0, float("inf")), _force_simt=False, _reuse_dt_cumsum=False, out=None):
    """Shard summary for the sequence-sharded path: (final state from a zero initial state (b,H,P,N) fp32,
    sum over the shard of dt*A (b,H) fp32).  `out` = (states, logdecay) buffers to fill (e.g. views of one
    flat all-gather send buffer).  New work -- SURVEY.md section 8e."""
    _, fin, logdecay = _ssd_call(x, dt, A, B, None, chunk_size, None, None, dt_bias, None, dt_softplus,
                                 dt_limit, L.TV_SSD_STATE_ONLY, force_simt=_force_simt, want_final=True,
                                 want_logdecay=True, reuse_dt_cumsum=_reuse_dt_cumsum,
                                 final_out=None if out is None else out[0],
                                 logdecay_out=None if out is None else out[1])
    return fin, logdecay


def mamba_dt_cumsum_prepare(x, dt, A, B, chunk_size, dt_bias=None, dt_softplus=False,
                            dt_limit=(0.0, float("inf")), workspace_stream=None):
    """Runs only the dt activation + per-chunk cumsum into the scratch owned by `workspace_stream` (default: the
    current stream), so that it can overlap the conv that produces x/B; x and B are NOT read (only their
    pointers/strides are inspected).  Follow with `_reuse_dt_cumsum=True` calls on the owning stream."""
    _ssd_call(x, dt, A, B, None, chunk_size, None, None, dt_bias, None, dt_softplus, dt_limit, L.TV_SSD_DT_ONLY,
              want_final=False, workspace_stream=workspace_stream)


def add_rmsnorm(x, weight, eps, residual=None):
    """Residual add + NemotronHRMSNorm in one pass (hybrid layer loop, modeling_nano.py:965 + :888-904):
    ``s = x + residual`` (rounded to the activation dtype, as the eager code materialises it) and
    ``out = (weight.float() * (s.float() * rsqrt(mean(s^2) + eps))).to(dtype)``.  Returns ``(out, s)``; with
    ``residual=None`` it is the plain norm of x and s is x."""
    _require_cuda(x, weight, residual)
    shape, d = x.shape, x.shape[-1]
    vec = 16 // x.element_size()

    def rows(t):
        t2 = t.reshape(-1, d)
        return t2 if (t2.stride(-1) == 1 and t2.stride(0) % vec == 0 and t2.data_ptr() % 16 == 0) else t2.contiguous()
    x2 = rows(x)
    r2 = None if residual is None else rows(residual.to(x.dtype))
    if residual is not None and residual.shape != shape:
        raise ValueError("add_rmsnorm: residual must have the shape of x")
    weight = weight.to(x.dtype).contiguous()
    out = torch.empty((x2.shape[0], d), dtype=x.dtype, device=x.device)
    summed = torch.empty_like(out) if r2 is not None else None
    if x2.shape[0] == 0:
        return out.reshape(shape), (x if residual is None else (x + residual))
    p = L.AddRmsnormParams(x=_ptr(x2), residual=_ptr(r2), weight=_ptr(weight), sum_out=_ptr(summed), out=_ptr(out),
                           rows=x2.shape[0], d=d, dtype=_dtype_code(x, "add_rmsnorm"), x_row_stride=x2.stride(0),
                           res_row_stride=0 if r2 is None else r2.stride(0),
                           sum_row_stride=0 if summed is None else summed.stride(0), out_row_stride=out.stride(0),
                           eps=float(eps), reserved=0)
    L.check(L.load().tv_add_rmsnorm_fwd(C.byref(p), _stream(x)), "add_rmsnorm")
    return out.reshape(shape), (x if summed is None else summed.reshape(shape))


def _rank_strided(t):
    """True if t (world, ...) is dense within each rank (only the rank stride may be padded)."""
    return t[0].is_contiguous() if t.shape[0] > 0 else True


def fold_boundary_states(states, logdecay, rank, initial_states=None):
    """states (world,b,H,P,N) fp32, logdecay (world,b,H) fp32 -> state entering shard `rank` (b,H,P,N).
    Both may be views of one flat gathered buffer (any rank stride)."""
    _require_cuda(states, logdecay, initial_states)
    world, b, H, P, N = states.shape
    states = states.to(torch.float32)
    logdecay = logdecay.to(torch.float32)
    if not _rank_strided(states) or states.stride(0) % 4 or states.data_ptr() % 16:
        states = states.contiguous()
    if not _rank_strided(logdecay):
        logdecay = logdecay.contiguous()
    init = None if initial_states is None else initial_states.to(torch.float32).contiguous()
    out = torch.empty((b, H, P, N), dtype=torch.float32, device=states.device)
    L.check(L.load().tv_ssd_fold_boundary_states(_ptr(states), _ptr(logdecay), _ptr(init), _ptr(out), int(rank),
                                                 b, H, P, N, states.stride(0) if world > 1 else 0,
                                                 logdecay.stride(0) if world > 1 else 0, _stream(states)),
            "fold_boundary_states")
    return out


def fold_boundary_states_p2p(state_ptrs, logdecay_ptrs, rank, shape, device, initial_states=None):
    """The fold of ``fold_boundary_states`` with the summary of rank r read through its own device pointer
    ``state_ptrs[r]`` / ``logdecay_ptrs[r]`` (ints; peer memory mapped into this process, r < rank): exchange and fold in
    one kernel over NVLink.  shape = (b, H, P, N).  Returns the state entering shard `rank`."""
    b, H, P, N = shape
    rank = int(rank)
    sp = (C.c_void_p * max(rank, 1))(*[C.c_void_p(int(v)) for v in state_ptrs[:rank]])
    lp = (C.c_void_p * max(rank, 1))(*[C.c_void_p(int(v)) for v in logdecay_ptrs[:rank]])
    init = None if initial_states is None else initial_states.to(torch.float32).contiguous()
    out = torch.empty((b, H, P, N), dtype=torch.float32, device=device)
    L.check(L.load().tv_ssd_fold_boundary_states_p2p(sp, lp, _ptr(init), _ptr(out), rank, b, H, P, N,
                                                     C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
            "fold_boundary_states_p2p")
    return out


def ssd_kernel_family(dtype, headdim, dstate, chunk_size, nheads=128, ngroups=8):
    """'tcgen05' or 'simt': which kernel family serves this shape (introspection for tests/bench)."""
    p = L.SsdParams(batch=1, seqlen=chunk_size, nheads=nheads, headdim=headdim, ngroups=ngroups, dstate=dstate,
                    chunk_size=chunk_size, dtype=_DT[dtype], mode=L.TV_SSD_FULL)
    return "tcgen05" if L.load().tv_ssd_kernel_family(C.byref(p)) == 1 else "simt"


# ------------------------------------------------------------------------------------------------
# single-token decode step (call sites modeling_nano.py:495-501, :528-539)
def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None, cache_seqlens=None, conv_state_indices=None):
    """x: (batch, dim) [or (batch, dim, 1)]; conv_state: (batch, dim, state_len >= width), updated IN PLACE (shifted left
    by one column, x appended); weight: (dim, width); bias: (dim,).  Returns out with the shape of x."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    if cache_seqlens is not None or conv_state_indices is not None:
        raise NotImplementedError("causal_conv1d_update: circular-buffer / indexed states are outside this path")
    _require_cuda(x, conv_state, weight, bias)
    squeeze = x.dim() == 3
    if squeeze:
        if x.shape[-1] != 1:
            raise NotImplementedError("causal_conv1d_update: one new token per call")
        x = x.squeeze(-1)
    b, dim = x.shape
    width = weight.shape[1]
    if conv_state.shape[:2] != (b, dim) or conv_state.shape[2] < width:
        raise ValueError("causal_conv1d_update: conv_state must be (batch, dim, state_len >= width)")
    if conv_state.dtype != x.dtype:
        raise ValueError("causal_conv1d_update: conv_state and x must have the same dtype")
    if conv_state.stride(2) != 1:        # e.g. a channel-last state: update a dense copy, then write it back in place
        dense = conv_state.contiguous()
        out = causal_conv1d_update(x, dense, weight, bias, activation)
        conv_state.copy_(dense)
        return out.unsqueeze(-1) if squeeze else out
    if x.stride(1) != 1:
        x = x.contiguous()
    weight = weight.to(x.dtype).contiguous()
    bias = None if bias is None else bias.to(x.dtype).contiguous()
    out = torch.empty((b, dim), dtype=x.dtype, device=x.device)
    p = L.ConvUpdateParams(x=_ptr(x), conv_state=_ptr(conv_state), weight=_ptr(weight), bias=_ptr(bias), out=_ptr(out),
                           batch=b, dim=dim, width=width, state_len=conv_state.shape[2], x_batch_stride=x.stride(0),
                           out_batch_stride=out.stride(0), state_batch_stride=conv_state.stride(0),
                           state_dim_stride=conv_state.stride(1), silu=int(activation in ("silu", "swish")),
                           dtype=_dtype_code(x, "causal_conv1d_update"))
    L.check(L.load().tv_causal_conv1d_update(C.byref(p), _stream(x)), "causal_conv1d_update")
    return out.unsqueeze(-1) if squeeze else out


def selective_state_update(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False,
                           state_batch_indices=None, _dt_limit=(0.0, float("inf"))):
    """state: (batch, nheads, headdim, dstate), updated IN PLACE; x, dt, z: (batch, nheads, headdim); A: (nheads,
    headdim, dstate); B, C: (batch, ngroups, dstate); D, dt_bias: (nheads, headdim).  Expanded (stride-0) views are
    taken as they are.  Returns out (batch, nheads, headdim)."""
    if state_batch_indices is not None:
        raise NotImplementedError("selective_state_update: indexed states are outside this path")
    if state.dim() != 4 or x.dim() != 3:
        raise NotImplementedError("selective_state_update: the multi-head form (state (b,H,P,N), x (b,H,P)) only")
    _require_cuda(state, x, dt, A, B, C, D, z, dt_bias)
    b, H, P, N = state.shape
    G = B.shape[1]
    assert x.shape == (b, H, P) and dt.shape == (b, H, P) and A.shape == (H, P, N)
    assert B.shape == (b, G, N) and C.shape == (b, G, N) and H % G == 0
    if D is not None:
        assert D.shape == (H, P)
    if z is not None:
        assert z.shape == x.shape
    if dt_bias is not None:
        assert dt_bias.shape == (H, P)
    if not state.is_contiguous():
        raise ValueError("selective_state_update: state must be contiguous (it is updated in place)")
    code = _dtype_code(x, "selective_state_update")
    dt, B, C = dt.to(x.dtype), B.to(x.dtype), C.to(x.dtype)
    if B.stride(-1) != 1:
        B = B.contiguous()
    if C.stride(-1) != 1:
        C = C.contiguous()
    z = None if z is None else z.to(x.dtype)
    f32 = lambda t: None if t is None else (t if t.dtype == torch.float32 else t.float())   # noqa: E731
    A32, D32, bias32 = f32(A), f32(D), f32(dt_bias)
    out = torch.empty((b, H, P), dtype=x.dtype, device=x.device)
    st3 = lambda t: (0, 0, 0) if t is None else tuple(t.stride())                            # noqa: E731
    st2 = lambda t: (0, 0) if t is None else tuple(t.stride())                               # noqa: E731
    lo, hi = float(_dt_limit[0]), float(_dt_limit[1])
    p = L.SsuParams(state=_ptr(state), x=_ptr(x), dt=_ptr(dt), A=_ptr(A32), B=_ptr(B), C=_ptr(C), D=_ptr(D32), z=_ptr(z),
                    dt_bias=_ptr(bias32), out=_ptr(out), batch=b, nheads=H, headdim=P, ngroups=G, dstate=N,
                    x_batch_stride=x.stride(0), x_head_stride=x.stride(1), x_dim_stride=x.stride(2),
                    dt_batch_stride=dt.stride(0), dt_head_stride=dt.stride(1), dt_dim_stride=dt.stride(2),
                    a_head_stride=A32.stride(0), a_dim_stride=A32.stride(1), a_state_stride=A32.stride(2),
                    b_batch_stride=B.stride(0), b_group_stride=B.stride(1),
                    c_batch_stride=C.stride(0), c_group_stride=C.stride(1),
                    d_head_stride=st2(D32)[0], d_dim_stride=st2(D32)[1],
                    z_batch_stride=st3(z)[0], z_head_stride=st3(z)[1], z_dim_stride=st3(z)[2],
                    bias_head_stride=st2(bias32)[0], bias_dim_stride=st2(bias32)[1],
                    dt_softplus=int(bool(dt_softplus)), dt_min=lo, dt_max=hi if math.isfinite(hi) else float("inf"),
                    dtype=code, state_dtype=_dtype_code(state, "selective_state_update(state)"))
    L.check(L.load().tv_selective_state_update(_byref(p), _stream(x)), "selective_state_update")
    return out


# -- the remaining name the reference's fast-path gate needs to be non-None (modeling_nano.py:89-97): training is
#    outside this path, so it fails loudly instead of silently computing something else -------------------------
def mamba_split_conv1d_scan_combined(*args, **kwargs):
    raise NotImplementedError("mamba_split_conv1d_scan_combined (training fwd+bwd) is outside the prefill path")
