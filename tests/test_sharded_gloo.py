"""Host-side logic of the sequence-sharded path (halo exchange, summary all-gather, fold order) with
world_size=2 and 4 over gloo on CPU.  The arithmetic is injected from the oracle; on the GPU box the same
code runs with the CUDA ops over NCCL (tests/test_gpu_sharded.py)."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mamba2_ref as R


def _oracle_ops():
    def conv(x, weight, bias=None, initial_states=None, activation=None, **kw):
        out, _ = R.causal_conv1d_ref(x, weight, bias, initial_states, activation, dtype=torch.float64)
        return out

    def summary(x, dt, A, B, chunk_size, dt_bias=None, dt_softplus=False, dt_limit=(0.0, float("inf")), **kw):
        C0 = torch.zeros_like(B)
        _, s = R.ssd_chunked_ref(x, dt, A, B, C0, chunk_size, dt_bias=dt_bias, dt_softplus=dt_softplus,
                                 dt_limit=dt_limit, dtype=torch.float64)
        logp = (R.dt_activate_ref(dt, dt_bias, dt_softplus, dt_limit, torch.float64) * A.double()).sum(1)
        return s, logp

    def fold(states, logdecay, rank, initial_states=None):
        return R.fold_boundary_states(list(states), list(logdecay), rank, initial_states)

    def scan(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None, dt_softplus=False,
             dt_limit=(0.0, float("inf")), return_final_states=False, **kw):
        y, s = R.ssd_chunked_ref(x, dt, A, B, C, chunk_size, D=D, z=z, dt_bias=dt_bias,
                                 initial_states=initial_states, dt_softplus=dt_softplus, dt_limit=dt_limit,
                                 dtype=torch.float64)
        return (y, s) if return_final_states else y

    def norm(x, weight, bias, z=None, eps=1e-6, group_size=None, norm_before_gate=True):
        return R.gated_rmsnorm_ref(x, weight, bias, z, eps, group_size, norm_before_gate, dtype=torch.float64)

    return types.SimpleNamespace(causal_conv1d_fn=conv, mamba_chunk_state_summary=summary,
                                 fold_boundary_states=fold, mamba_chunk_scan_combined=scan, rmsnorm_fn=norm)


def _worker(rank, world, port, L, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import timeviper_b200 as tv
        torch.manual_seed(7)
        cfg = tv.Mamba2Config(hidden_size=32, mamba_num_heads=4, mamba_head_dim=8, n_groups=2, ssm_state_size=16,
                              chunk_size=32)
        mixer = tv.Mamba2MixerPrefill(cfg).double()
        with torch.no_grad():
            mixer.A_log.copy_(torch.log(torch.rand(4) * 3 + 0.05))
            mixer.dt_bias.copy_(torch.randn(4) * 0.5 - 2.0)
            mixer.D.copy_(torch.randn(4))
        hs = torch.randn(1, L, 32, dtype=torch.float64)
        with torch.no_grad():
            proj = mixer.in_proj(hs)
            # unsharded truth with the same oracle ops
            p = {k: v.detach() for k, v in mixer.state_dict().items()}
            ref_out, ref_conv, ref_ssm = R.mixer_forward_ref(
                p, hs, num_heads=4, head_dim=8, n_groups=2, ssm_state_size=16, chunk_size=32,
                dtype=torch.float64)
            sl = slice(rank * L // world, (rank + 1) * L // world)
            cache = types.SimpleNamespace(conv_kernel_size=4, conv=None, ssm=None)
            cache.update_conv_state = lambda layer_idx, new_conv_state, cache_init: setattr(cache, "conv", new_conv_state)
            cache.update_ssm_state = lambda layer_idx, new_ssm_state: setattr(cache, "ssm", new_ssm_state)
            out = tv.sharded_mixer_forward(mixer, hs[:, sl], cache_params=cache, ops=_oracle_ops())
        err = float((out - ref_out[:, sl]).abs().max() / ref_out.abs().max())
        res = {"rank": rank, "err": err}
        if rank == world - 1:
            res["ssm_err"] = float((cache.ssm - ref_ssm).abs().max() / ref_ssm.abs().max())
            res["conv_equal"] = bool(torch.equal(cache.conv, ref_conv))
        else:
            res["cache_untouched"] = cache.ssm is None and cache.conv is None
        out_q.put(res)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,L", [(2, 128), (4, 256), (2, 70), (8, 512)])
def test_sharded_equals_unsharded_gloo(world, L):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert r["err"] < 1e-11, r
        if r["rank"] == world - 1:
            assert r["ssm_err"] < 1e-11 and r["conv_equal"], r
        else:
            assert r["cache_untouched"], r


def _hybrid_worker(rank, world, port, L, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import timeviper_b200 as tv
        torch.manual_seed(11)
        pattern = "M*-M*"
        cfg = tv.Mamba2Config(hidden_size=32, mamba_num_heads=4, mamba_head_dim=8, n_groups=2, ssm_state_size=16,
                              chunk_size=32, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                              num_attention_heads=4, num_key_value_heads=2, head_dim=8, intermediate_size_mlp=48, vocab_size=50)
        model = tv.HybridCausalLM(cfg).double()
        with torch.no_grad():
            for layer in model.backbone.layers:
                if layer.block_type == "mamba":
                    layer.mixer.A_log.copy_(torch.log(torch.rand(4) * 3 + 0.05))
                    layer.mixer.dt_bias.copy_(torch.randn(4) * 0.5 - 2.0)
                    layer.mixer.D.copy_(torch.randn(4))
        x = torch.randn(1, L, 32, dtype=torch.float64)
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        bb = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
        ref_h = R.hybrid_forward_ref(bb, x, pattern=pattern, num_heads=4, head_dim=8, n_groups=2, ssm_state_size=16,
                                     chunk_size=32, attn_heads=4, kv_heads=2, attn_head_dim=8, dtype=torch.float64)
        ref_logits = torch.nn.functional.linear(ref_h[:, -1:], sd["lm_head.weight"])
        sl = slice(rank * L // world, (rank + 1) * L // world)
        with torch.no_grad():
            h = model.backbone(inputs_embeds=x[:, sl], group=dist.group.WORLD, mixer_ops=_oracle_ops())
            logits = model.lm_head(h[:, -1:])
        res = {"rank": rank, "err": float((h - ref_h[:, sl]).abs().max() / ref_h.abs().max())}
        if rank == world - 1:
            res["logit_err"] = float((logits - ref_logits).abs().max() / ref_logits.abs().max())
        out_q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L", [(2, 128), (4, 192), (3, 99)])
def test_sharded_hybrid_stack_equals_unsharded_gloo(world, L):
    """Sequence-sharded hybrid layer loop (Mamba-2 layers: halo + boundary states; attention layers: K/V all-gather with a
    lower-right causal mask; MLP / norms token-local) against the unsharded oracle stack, fp64, gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, world, port, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:       # the stack's RMSNorm takes its statistics in fp32 (as the reference's does): 1e-7, not 1e-11
        assert r["err"] < 1e-6, r
        if r["rank"] == world - 1:
            assert r["logit_err"] < 1e-6, r


def _pdrop_worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import timeviper_b200 as tv
        from timeviper_b200.hybrid import shard_bounds
        torch.manual_seed(23)
        pattern, pre, V, post = "M*-M*", 3, 100, 10
        cfg = tv.Mamba2Config(hidden_size=32, mamba_num_heads=4, mamba_head_dim=8, n_groups=2, ssm_state_size=16,
                              chunk_size=32, num_hidden_layers=len(pattern), hybrid_override_pattern=pattern,
                              num_attention_heads=4, num_key_value_heads=2, head_dim=8, intermediate_size_mlp=48, vocab_size=50)
        model = tv.HybridPrefillStack(cfg).double()
        with torch.no_grad():
            for layer in model.layers:
                if layer.block_type == "mamba":
                    layer.mixer.A_log.copy_(torch.log(torch.rand(4) * 3 + 0.05))
                    layer.mixer.dt_bias.copy_(torch.randn(4) * 0.5 - 2.0)
                    layer.mixer.D.copy_(torch.randn(4))
        L = pre + V + post
        x = torch.randn(1, L, 32, dtype=torch.float64)
        pd = dict(pdrop_type="uni_2_0.8-attn_4_0.5", first_vision_token_position=pre, num_vision_tokens=V, text_prompt_len=pre + post)
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        ref = R.hybrid_forward_ref(sd, x, pattern=pattern, num_heads=4, head_dim=8, n_groups=2, ssm_state_size=16,
                                   chunk_size=32, attn_heads=4, kv_heads=2, attn_head_dim=8, dtype=torch.float64, pdrop=pd)
        offs = shard_bounds(L, world)
        with torch.no_grad():
            h = model(inputs_embeds=x[:, offs[rank]:offs[rank + 1]], group=dist.group.WORLD, mixer_ops=_oracle_ops(), pdrop=pd)
        new = shard_bounds(ref.shape[1], world)
        mine = ref[:, new[rank]:new[rank + 1]]
        out_q.put({"rank": rank, "shape_ok": tuple(h.shape) == tuple(mine.shape),
                   "err": float((h - mine).abs().max() / ref.abs().max()) if tuple(h.shape) == tuple(mine.shape) else 1.0})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pyramid_drop_equals_unsharded_gloo(world):
    """TransV / pyramid-drop over a sequence-sharded sample (uniform stage, then an attention-ranked stage with the
    distributed softmax and the all-to-all re-balancing; unequal shards) against the unsharded oracle stack, fp64, gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pdrop_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert r["shape_ok"] and r["err"] < 1e-6, r
