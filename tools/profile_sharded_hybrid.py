"""torchrun helper: one Mamba-2 block of the 9B shape, sequence-sharded vs the same tokens unsharded per rank
(time per call with CUDA events; rank 0 prints the profiler table of the sharded call).
    torchrun --nproc-per-node 2 tools/profile_sharded_hybrid.py [tokens_per_rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import timeviper_b200 as tv
from timeviper_b200.hybrid import HybridBlock

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
cfg = tv.Mamba2Config(num_hidden_layers=1, hybrid_override_pattern="M")
torch.manual_seed(0)
with torch.device("cuda"):
    blk = HybridBlock(cfg, 0).to(torch.bfloat16).eval()
h = torch.randn(1, L, cfg.hidden_size, device="cuda").to(torch.bfloat16)
pos = torch.arange(L)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_un = timeit(lambda: blk(h, cache_position=pos))
    t_sh = timeit(lambda: blk(h, cache_position=pos, group=dist.group.WORLD))
    print(f"rank {rank}: {L} tokens/rank, unsharded block {t_un:.2f} ms, sharded block {t_sh:.2f} ms", flush=True)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            blk(h, cache_position=pos, group=dist.group.WORLD)
        torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
dist.destroy_process_group()
