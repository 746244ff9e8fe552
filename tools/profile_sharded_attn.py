"""torchrun helper: the sequence-sharded attention layer of the 9B shape, per rank, under each SDPA backend.
    torchrun --nproc-per-node 2 tools/profile_sharded_attn.py [tokens_per_rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.nn.attention import sdpa_kernel, SDPBackend
import timeviper_b200 as tv
from timeviper_b200.hybrid import Attention, sharded_attention_forward

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
cfg = tv.Mamba2Config()
torch.manual_seed(0)
with torch.device("cuda"):
    attn = Attention(cfg, 0).to(torch.bfloat16).eval()
h = torch.randn(1, L, cfg.hidden_size, device="cuda").to(torch.bfloat16)


def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    print(f"rank {rank}: unsharded attention on {L} tokens {timeit(lambda: attn(h)):.2f} ms", flush=True)
    cases = (("default", None),)
    if os.environ.get("TV_ALL_BACKENDS"):     # a backend that rejects the call on one rank only leaves the other in a collective
        cases += (("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION))
    for name, be in cases:
        try:
            if be is None:
                t = timeit(lambda: sharded_attention_forward(attn, h, dist.group.WORLD))
            else:
                with sdpa_kernel(be):
                    t = timeit(lambda: sharded_attention_forward(attn, h, dist.group.WORLD))
            print(f"rank {rank}: sharded attention [{name}] {t:.2f} ms", flush=True)
        except Exception as e:
            print(f"rank {rank}: sharded attention [{name}] failed: {str(e)[:120]}", flush=True)
            torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
