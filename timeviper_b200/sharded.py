"""Sequence-sharded Mamba-2 mixer prefill across the GPUs of one node (one process per GPU).

New work: the reference has no sequence/context parallelism (SURVEY.md sections 5, 8e).  Rank r holds the
contiguous shard r of the sequence (all heads, replicated parameters).  Per layer:

  1. local: in_proj.  The K-1 = 3 pre-conv rows that precede the shard (conv halo) are all-gathered on a HELPER
     stream while the main stream already runs the conv over the whole shard with a zero halo; the helper stream
     also runs the dt softplus/cumsum (it needs only dt) and, once the halo is there, re-does the conv of the first
     K-1 rows (the only ones that see the halo).  Join; patch the K-1 rows.
  2. local pass 1: shard summary (S_r = state after the shard from a zero state, log P_r = sum dt*A), written
     straight into one flat send buffer [S_r | log P_r].
  3. ONE collective on the critical path: all-gather of that buffer -- 5.24 MB + 512 B per rank at the 9B dims,
     independent of L -- over NCCL/NVLink.
  4. local: fold ranks < r in fp32:  S_in(r+1) = exp(log P_r) S_in(r) + S_r   (reads the gathered buffer in place).
  5. local pass 2: the full scan with ``initial_states = S_in(r)`` (dt/cumsum reused); gated norm; out_proj.

The final SSM state of the sequence is rank W-1's; the final conv state is rank W-1's last K rows.
``ops`` is injectable so that the host-side logic (halo patch, gather, fold order) is testable with gloo on CPU.
"""
import contextlib

import torch
import torch.distributed as dist
from torch import nn

from . import ops as _cuda_ops

_helper_streams = {}


def _helper_stream(device):
    """High-priority side stream: its small kernels (dt cumsum, halo conv) are scheduled into the SM slots that the
    shard-wide conv on the main stream frees, instead of queueing behind its whole grid."""
    s = _helper_streams.get(device.index)
    if s is None:
        s = _helper_streams[device.index] = torch.cuda.Stream(device, priority=-1)
    return s


def _all_gather_rows(send, group):
    """send: contiguous tensor -> (world, *send.shape), one collective, no staging copy."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    if send.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), send.view(-1), group=group)
    else:                       # gloo: list form, the outputs are views of `out`
        dist.all_gather(list(out.unbind(0)), send, group=group)
    return out


class PeerExchange:
    """Symmetric-memory buffers of the two exchanges of a sharded layer (one object per (group, device, shapes)):

      * `flat`  [2][S_r | log P_r] fp32: the shard summary is written by the state kernel straight into this rank's
        buffer; after one device-side barrier every rank folds the summaries it needs by reading its peers' buffers
        over NVLink inside the fold kernel (ops.fold_boundary_states_p2p) -- no all-gather, no gathered copy;
      * `halo`  [2][b, K-1, conv_dim]: rank r stores the last K-1 pre-conv rows of its shard into rank r+1's buffer
        (one small peer copy) and signals it.

    Both are double-buffered by step parity: a rank that is a step ahead writes the other half, and it cannot be two
    steps ahead because every step ends with the barrier."""

    def __init__(self, group, device, b, H, P, N, halo_shape, halo_dtype):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = b * H * P * N
        self.per = self.n + b * H
        self.per_al = (self.per + 63) // 64 * 64                       # keep the second half 256-byte aligned
        self.flat = symm.empty(2 * self.per_al, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.flat, self.group)
        self.halo_elems = 1
        for v in halo_shape:
            self.halo_elems *= v
        self.halo_al = (self.halo_elems + 127) // 128 * 128
        self.halo = symm.empty(2 * self.halo_al, dtype=halo_dtype, device=device)
        self.halo_hdl = symm.rendezvous(self.halo, self.group)
        self.halo_shape, self.shape = tuple(halo_shape), (b, H, P, N)
        self.last_parity = 1                          # the first step uses half 0
        self.forced_parity = None                     # set while a CUDA graph of a given parity is captured

    def next_parity(self):
        """Half of the double buffers this step uses: always the other one than the previous step (eager or replayed)."""
        p = self.forced_parity if self.forced_parity is not None else self.last_parity ^ 1
        self.last_parity = p
        return p

    def summary_views(self, parity):
        base = self.flat[parity * self.per_al: parity * self.per_al + self.per]
        b, H, P, N = self.shape
        return base[:self.n].view(b, H, P, N), base[self.n:].view(b, H)

    def peer_ptrs(self, parity):
        off = parity * self.per_al * 4
        sp = [int(p) + off for p in self.hdl.buffer_ptrs]
        return sp, [p + self.n * 4 for p in sp]

    def send_halo(self, rows, parity):
        """rows (b, K-1, conv_dim) -> the halo buffer of rank+1, then signal it (stream-ordered)."""
        if self.rank + 1 < self.world:
            dst = self.halo_hdl.get_buffer(self.rank + 1, self.halo_shape, self.halo.dtype, parity * self.halo_al)
            dst.copy_(rows)
            self.halo_hdl.put_signal(self.rank + 1, channel=parity)

    def recv_halo(self, parity):
        """Wait for rank-1's rows; returns the local (b, K-1, conv_dim) buffer (None on rank 0)."""
        if self.rank == 0:
            return None
        self.halo_hdl.wait_signal(self.rank - 1, channel=parity)
        return self.halo[parity * self.halo_al: parity * self.halo_al + self.halo_elems].view(self.halo_shape)


_exchanges = {}
_p2p_disabled = [False]


def _peer_exchange(group, dev, b, H, P, N, halo_shape, halo_dtype):
    """The PeerExchange for these shapes, or None where symmetric memory is not usable (then NCCL all-gathers are used)."""
    import os
    if _p2p_disabled[0] or os.environ.get("TV_SHARDED_EXCHANGE", "p2p").lower() == "nccl":
        return None
    key = (id(group), dev.index, b, H, P, N, tuple(halo_shape), halo_dtype)
    ex = _exchanges.get(key)
    if ex is None:
        try:
            ex = PeerExchange(group, dev, b, H, P, N, halo_shape, halo_dtype)
        except Exception as e:             # no P2P / fabric support on this box: keep the collective transport
            import warnings
            warnings.warn(f"timeviper_b200: symmetric-memory exchange unavailable ({type(e).__name__}: {e}); "
                          "falling back to NCCL all-gathers for the boundary exchange")
            _p2p_disabled[0] = True
            return None
        _exchanges[key] = ex
    return ex


def _gathered_fold(ops, native, mixer, x, dt, A, B, scan_kw, b, H, P, N, n, xBC_c, dev, group, rank, world):
    """Collective transport of the boundary exchange (NCCL / gloo): summary into one flat buffer, ONE all-gather, fold."""
    flat = torch.empty(n + b * H, dtype=torch.float32 if native else xBC_c.dtype, device=dev)
    S_r, logP_r = flat[:n].view(b, H, P, N), flat[n:].view(b, H)
    if native:

        ops.mamba_chunk_state_summary(x, dt, A, B, mixer.chunk_size, _reuse_dt_cumsum=True, out=(S_r, logP_r),
                                      **scan_kw)
    else:
        s, lp = ops.mamba_chunk_state_summary(x, dt, A, B, mixer.chunk_size, **scan_kw)
        S_r.copy_(s)
        logP_r.copy_(lp)
    gathered = _all_gather_rows(flat, group)                          # (world, n + b*H)
    S_all = gathered[:, :n].view(world, b, H, P, N)
    logP_all = gathered[:, n:].view(world, b, H)
    S_in = ops.fold_boundary_states(S_all, logP_all, rank) if rank > 0 else None
    return S_in


def sharded_scan_core(mixer, projected_states, group=None, cache_params=None, ops=_cuda_ops, attention_mask=None):
    """The three-kernel core of ``Mamba2MixerPrefill.scan_core`` on this rank's shard of the sequence."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, L, _ = projected_states.shape
    if attention_mask is not None and b > 1 and L > 1:
        raise NotImplementedError("the sequence-sharded path takes unpadded sequences (the reference is batch-1 here)")
    K = mixer.conv_kernel_size
    H, P, G, N = mixer.num_heads, mixer.head_dim, mixer.n_groups, mixer.ssm_state_size
    gate, xBC, dt = projected_states.split([mixer.intermediate_size, mixer.conv_dim, H], dim=-1)
    assert L >= K - 1, "a shard must hold at least conv_kernel-1 tokens"
    native = ops is _cuda_ops
    dev = projected_states.device
    A = mixer.decay_rates().to(dev)
    w, bias = mixer.conv1d.weight.squeeze(1), mixer.conv1d.bias
    scan_kw = dict(dt_bias=mixer.f32_param("dt_bias"), dt_softplus=True, dt_limit=mixer.time_step_limit)

    # 1. conv over the shard (zero halo) || halo exchange, dt cumsum, conv of the first K-1 rows with the halo
    halo_send = xBC[:, L - (K - 1):, :].contiguous()                  # (b, K-1, conv_dim)
    ex = _peer_exchange(group, dev, b, H, P, N, (b, K - 1, mixer.conv_dim), xBC.dtype) if (native and world <= 16) else None
    parity = ex.next_parity() if ex is not None else 0
    if native:
        main, side = torch.cuda.current_stream(dev), _helper_stream(dev)
        xBC_c = torch.empty((b, L, mixer.conv_dim), dtype=xBC.dtype, device=dev)
        side.wait_stream(main)
        side_ctx = torch.cuda.stream(side)
        # the big kernel is enqueued first; the helper stream's small ones overtake it on the device (priority)
        ops.causal_conv1d_into(xBC_c, x=xBC.transpose(1, 2), weight=w, bias=bias, activation=mixer.activation)
    else:
        xBC_c, side_ctx = None, contextlib.nullcontext()

    def views(t):
        x, B, C = torch.split(t, [mixer.intermediate_size, G * N, G * N], dim=-1)
        return x.view(b, L, H, P), B.view(b, L, G, N), C.view(b, L, G, N)

    head_rows = None
    with side_ctx:
        if ex is not None:                                            # peer store into rank+1's buffer + signal
            ex.send_halo(halo_send, parity)
            halos = None
        else:
            halos = _all_gather_rows(halo_send, group)                # (world, b, K-1, conv_dim)
        if native:                                                    # needs dt only; scratch of the main stream
            x, B, C = views(xBC_c)
            ops.mamba_dt_cumsum_prepare(x, dt, A, B, mixer.chunk_size, workspace_stream=main, **scan_kw)
        prev_halo = None
        if rank > 0:
            prev_halo = ex.recv_halo(parity) if ex is not None else halos[rank - 1]
            conv_init = prev_halo.transpose(1, 2).contiguous()        # (b, conv_dim, K-1)
            head_rows = ops.causal_conv1d_fn(x=xBC[:, :K - 1].transpose(1, 2), weight=w, bias=bias,
                                             initial_states=conv_init, activation=mixer.activation)
    if native:
        main.wait_stream(side)
        for t in (halos, head_rows, prev_halo if ex is None else None):
            if t is not None:
                t.record_stream(main)
    else:
        xBC_c = ops.causal_conv1d_fn(x=xBC.transpose(1, 2), weight=w, bias=bias,
                                     activation=mixer.activation).transpose(1, 2).contiguous()
    if head_rows is not None:
        xBC_c[:, :K - 1].copy_(head_rows.transpose(1, 2))
    x, B, C = views(xBC_c)

    # 2.-4. shard summary, exchange, fold
    n = b * H * P * N
    if ex is not None:
        # summary straight into this rank's symmetric buffer; ONE device-side barrier; every rank then reads the summaries
        # it needs from its peers inside the fold kernel (NVLink loads) -- no all-gather on the critical path
        S_r, logP_r = ex.summary_views(parity)
        ops.mamba_chunk_state_summary(x, dt, A, B, mixer.chunk_size, _reuse_dt_cumsum=True, out=(S_r, logP_r), **scan_kw)
        ex.hdl.barrier(channel=parity)
        if rank > 0:
            sp, lp = ex.peer_ptrs(parity)
            S_in = ops.fold_boundary_states_p2p(sp, lp, rank, (b, H, P, N), dev)
        else:
            S_in = None
    else:
        S_in = _gathered_fold(ops, native, mixer, x, dt, A, B, scan_kw, b, H, P, N, n, xBC_c, dev, group, rank, world)

    # 5. full local scan from the folded entering state (dt/cumsum of step 1 is still in the op's scratch)
    reuse = {"_reuse_dt_cumsum": True} if native else {}
    y, ssm_state = ops.mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size=mixer.chunk_size, D=mixer.f32_param("D"), z=None,
                                                 initial_states=S_in, return_final_states=True, **scan_kw, **reuse)
    if cache_params is not None and rank == world - 1:
        xt = xBC.transpose(1, 2)
        conv_states = nn.functional.pad(xt, (cache_params.conv_kernel_size - xt.shape[-1], 0))
        if L < K:   # left columns come from the previous shard, not zeros
            conv_states[..., :K - L] = prev_halo.transpose(1, 2)[..., L - K:] if rank > 0 else 0
        cache_params.update_conv_state(layer_idx=mixer.layer_idx, new_conv_state=conv_states, cache_init=True)
        cache_params.update_ssm_state(layer_idx=mixer.layer_idx, new_ssm_state=ssm_state)
    y = ops.rmsnorm_fn(x=y.view(b, L, -1), weight=mixer.norm.weight, bias=None, z=gate,
                       eps=mixer.norm.variance_epsilon, group_size=mixer.norm.group_size, norm_before_gate=False)
    return y, ssm_state


_sharded_graphs = {}


@torch.no_grad()
def sharded_scan_core_graph(mixer, projected_states, group=None):
    """``sharded_scan_core`` replayed as ONE CUDA graph launch per step.  Possible because the boundary exchange is made of
    ordinary kernels and peer copies (symmetric memory), not of NCCL calls.  Two graphs are captured, one per half of the
    exchange's double buffers, and replayed alternately; the returned tensors are the graph's static outputs (overwritten
    by the replay after next).  Falls back to eager launches where the peer exchange is unavailable."""
    dev = projected_states.device
    key = (projected_states.data_ptr(), tuple(projected_states.shape), projected_states.dtype, id(group), dev.index)
    held = _sharded_graphs.get(key)
    if held is None:
        for _ in range(2):                                # both halves once, eagerly: allocations, lazy CUDA state
            sharded_scan_core(mixer, projected_states, group=group)
        ex = next((e for k, e in _exchanges.items() if k[1] == dev.index), None)
        if ex is None:
            held = _sharded_graphs[key] = {"graphs": None}
        else:
            torch.cuda.synchronize(dev)
            dist.barrier(group)
            graphs = {}
            for parity in (0, 1):
                ex.forced_parity = parity
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    y, ssm = sharded_scan_core(mixer, projected_states, group=group)
                graphs[parity] = (g, y, ssm)
            ex.forced_parity = None
            torch.cuda.synchronize(dev)
            dist.barrier(group)
            held = _sharded_graphs[key] = {"graphs": graphs, "ex": ex}
    if held["graphs"] is None:
        return sharded_scan_core(mixer, projected_states, group=group)
    parity = held["ex"].next_parity()
    g, y, ssm = held["graphs"][parity]
    g.replay()
    return y, ssm


def sharded_mixer_forward(mixer, hidden_states_shard, group=None, cache_params=None, ops=_cuda_ops):
    """hidden_states_shard: this rank's (b, L/W, hidden) slice.  Returns this rank's (b, L/W, hidden) output."""
    projected = mixer.in_proj(hidden_states_shard)
    y, _ = sharded_scan_core(mixer, projected, group=group, cache_params=cache_params, ops=ops)
    return mixer.out_proj(y)


class _HostPipe:
    """Per-device double buffers, copy streams and events of ``sharded_prefill_from_host``."""

    def __init__(self, dev, in_shape, out_shape, dtype):
        self.key = (tuple(in_shape), tuple(out_shape), dtype)
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.d_in = [torch.empty(in_shape, dtype=dtype, device=dev) for _ in range(2)]
        self.y = [None, None]                       # keeps the output of slot i alive until its D2H copy has run
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_compute = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.step = 0


_host_pipes = {}


@torch.no_grad()
def sharded_prefill_from_host(mixer, hidden_host_shard, out_host_shard=None, group=None, cache_params=None):
    """This rank's shard from (pinned) HOST memory to (pinned) host memory.  The H2D copy, the sharded mixer and the D2H
    copy run on three streams with double-buffered device tensors, so when shards arrive back to back (one call per
    sequence) the H2D copy of the next call overlaps the D2H copy of the previous one -- the two PCIe directions are
    independent, and within ONE sequence they cannot overlap because every output depends on every earlier token.
    The call returns once the work is enqueued; ``torch.cuda.synchronize()`` (or the returned event) before reading
    ``out_host_shard``.  Returns (out_host_shard, done_event)."""
    dev = mixer.in_proj.weight.device
    b, L, _ = hidden_host_shard.shape
    out_shape = (b, L, mixer.hidden_size)
    if out_host_shard is None:
        out_host_shard = torch.empty(out_shape, dtype=hidden_host_shard.dtype).pin_memory()
    pipe = _host_pipes.get(dev.index)
    if pipe is None or pipe.key != (tuple(hidden_host_shard.shape), out_shape, hidden_host_shard.dtype):
        torch.cuda.synchronize(dev)
        pipe = _host_pipes[dev.index] = _HostPipe(dev, hidden_host_shard.shape, out_shape, hidden_host_shard.dtype)
    i = pipe.step % 2
    first_use = pipe.step < 2
    pipe.step += 1
    cur = torch.cuda.current_stream(dev)
    with torch.cuda.stream(pipe.s_in):
        if not first_use:
            pipe.s_in.wait_event(pipe.ev_compute[i])          # the mixer of two calls ago has read this input buffer
        else:
            pipe.s_in.wait_stream(cur)
        pipe.d_in[i].copy_(hidden_host_shard, non_blocking=True)
        pipe.ev_in[i].record(pipe.s_in)
    cur.wait_event(pipe.ev_in[i])
    y = sharded_mixer_forward(mixer, pipe.d_in[i], group=group, cache_params=cache_params)
    pipe.ev_compute[i].record(cur)
    with torch.cuda.stream(pipe.s_out):
        pipe.s_out.wait_event(pipe.ev_compute[i])
        out_host_shard.copy_(y, non_blocking=True)
        pipe.ev_out[i].record(pipe.s_out)
    y.record_stream(pipe.s_out)
    pipe.y[i] = y
    return out_host_shard, pipe.ev_out[i]
