"""sharded_prefill_from_host (double-buffered H2D / sharded mixer / D2H) equals the unsharded mixer, over several
back-to-back calls with different inputs (buffer reuse).  Needs 2 GPUs; `pytest -m gpu`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import timeviper_b200 as tv
        from oracle import mamba2_ref as R
        cfg = tv.Mamba2Config(hidden_size=256, mamba_num_heads=16, mamba_head_dim=80, n_groups=2, ssm_state_size=128,
                              chunk_size=128)
        p = R.nemotron_random_params(cfg.hidden_size, 16, 80, 2, 128)
        mixer = tv.Mamba2MixerPrefill(cfg).to(torch.bfloat16).cuda()
        mixer.load_state_dict({k: v.to(torch.bfloat16) for k, v in p.items()}, strict=True)
        L = 1024
        errs = []
        outs = []
        with torch.no_grad():
            for step in range(4):
                g = torch.Generator().manual_seed(100 + step)
                hs = torch.randn(1, L, cfg.hidden_size, generator=g).to(torch.bfloat16)
                sl = slice(rank * L // world, (rank + 1) * L // world)
                host_in = hs[:, sl].contiguous().pin_memory()
                host_out = torch.empty(1, L // world, cfg.hidden_size, dtype=torch.bfloat16).pin_memory()
                outs.append((hs, sl, host_out, tv.sharded_prefill_from_host(mixer, host_in, host_out)[1]))
            torch.cuda.synchronize()
            for hs, sl, host_out, _ in outs:
                ref = mixer(hs.cuda())[:, sl].cpu()
                errs.append(float((host_out.float() - ref.float()).abs().max() / ref.float().abs().max()))
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


def test_sharded_prefill_from_host_equals_unsharded():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 2e-2, (rank, errs)
