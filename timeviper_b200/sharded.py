"""Sequence-sharded Mamba-2 mixer prefill across the GPUs of one node (one process per GPU).

New work: the reference has no sequence/context parallelism (SURVEY.md sections 5, 8e).  Rank r holds the
contiguous shard r of the sequence (all heads, replicated parameters).  Per layer:

  1. local: in_proj; exchange the K-1 = 3 pre-conv rows that precede the shard (conv halo) with an
     all-gather; conv with ``initial_states`` = halo of rank r-1.
  2. local pass 1: shard summary (S_r = state after the shard from a zero state, log P_r = sum dt*A).
  3. ONE collective: all-gather of (S_r, log P_r)  -- 5.24 MB + 512 B per rank at the 9B dims,
     independent of L -- over NCCL/NVLink.
  4. local: fold ranks < r in fp32:  S_in(r+1) = exp(log P_r) S_in(r) + S_r.
  5. local pass 2: the full scan with ``initial_states = S_in(r)``; gated norm; out_proj.

The final SSM state of the sequence is rank W-1's; the final conv state is rank W-1's last K rows.
``ops`` is injectable so that the host-side logic (halo, gather, fold order) is testable with gloo on CPU.
"""
import torch
import torch.distributed as dist
from torch import nn

from . import ops as _cuda_ops


def _all_gather_cat(t, group):
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather(list(out.unbind(0)), t.contiguous(), group=group)   # views of `out`: no extra copy
    return out


def sharded_scan_core(mixer, projected_states, group=None, cache_params=None, ops=_cuda_ops):
    """The three-kernel core of ``Mamba2MixerPrefill.scan_core`` on this rank's shard of the sequence."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, L, _ = projected_states.shape
    K = mixer.conv_kernel_size
    gts = mixer.n_groups * mixer.ssm_state_size
    gate, xBC, dt = projected_states.split([mixer.intermediate_size, mixer.conv_dim, mixer.num_heads], dim=-1)

    # 1. conv halo: last K-1 pre-conv rows of every shard (b, K-1, conv_dim)
    assert L >= K - 1, "a shard must hold at least conv_kernel-1 tokens"
    halos = _all_gather_cat(xBC[:, L - (K - 1):, :], group)
    conv_init = None if rank == 0 else halos[rank - 1].transpose(1, 2).contiguous()   # (b, conv_dim, K-1)
    xBC_c = ops.causal_conv1d_fn(x=xBC.transpose(1, 2), weight=mixer.conv1d.weight.squeeze(1),
                                 bias=mixer.conv1d.bias, initial_states=conv_init,
                                 activation=mixer.activation).transpose(1, 2)
    x, B, C = torch.split(xBC_c, [mixer.intermediate_size, gts, gts], dim=-1)
    x = x.view(b, L, -1, mixer.head_dim)
    B = B.view(b, L, mixer.n_groups, -1)
    C = C.view(b, L, mixer.n_groups, -1)
    A = -torch.exp(mixer.A_log.float())

    # 2.-4. shard summary, one all-gather, local fold
    S_r, logP_r = ops.mamba_chunk_state_summary(x, dt, A, B, mixer.chunk_size, dt_bias=mixer.dt_bias,
                                                dt_softplus=True, dt_limit=mixer.time_step_limit)
    packed = torch.cat([S_r.reshape(b, mixer.num_heads, -1), logP_r[..., None]], dim=-1)   # (b,H,P*N+1)
    gathered = _all_gather_cat(packed, group)
    S_all = gathered[..., :-1].reshape(world, b, mixer.num_heads, mixer.head_dim, mixer.ssm_state_size)
    logP_all = gathered[..., -1]
    S_in = ops.fold_boundary_states(S_all, logP_all, rank) if rank > 0 else None

    # 5. full local scan from the folded entering state (dt/cumsum of pass 1 is still in the op's workspace)
    reuse = {"_reuse_dt_cumsum": True} if ops is _cuda_ops else {}
    y, ssm_state = ops.mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size=mixer.chunk_size, D=mixer.D, z=None,
                                                 dt_bias=mixer.dt_bias, dt_softplus=True,
                                                 dt_limit=mixer.time_step_limit, initial_states=S_in,
                                                 return_final_states=True, **reuse)
    if cache_params is not None and rank == world - 1:
        xt = xBC.transpose(1, 2)
        conv_states = nn.functional.pad(xt, (cache_params.conv_kernel_size - xt.shape[-1], 0))
        if L < K:   # left columns come from the previous shard, not zeros
            conv_states[..., :K - L] = halos[rank - 1].transpose(1, 2)[..., L - K:] if rank > 0 else 0
        cache_params.update_conv_state(layer_idx=mixer.layer_idx, new_conv_state=conv_states, cache_init=True)
        cache_params.update_ssm_state(layer_idx=mixer.layer_idx, new_ssm_state=ssm_state)
    y = ops.rmsnorm_fn(x=y.view(b, L, -1), weight=mixer.norm.weight, bias=None, z=gate,
                       eps=mixer.norm.variance_epsilon, group_size=mixer.norm.group_size, norm_before_gate=False)
    return y, ssm_state


def sharded_mixer_forward(mixer, hidden_states_shard, group=None, cache_params=None, ops=_cuda_ops):
    """hidden_states_shard: this rank's (b, L/W, hidden) slice.  Returns this rank's (b, L/W, hidden) output."""
    projected = mixer.in_proj(hidden_states_shard)
    y, _ = sharded_scan_core(mixer, projected, group=group, cache_params=cache_params, ops=ops)
    return mixer.out_proj(y)
