"""torchrun helper: per-kernel CUDA time of one sharded scan_core step (rank 0 prints the table)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import timeviper_b200 as tv
from oracle import mamba2_ref as R

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                        pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
cfg = tv.Mamba2Config.nanov2_9b()
L = int(sys.argv[1]) // world if len(sys.argv) > 1 else 131072 // world
p = R.nemotron_random_params(cfg.hidden_size, cfg.mamba_num_heads, cfg.mamba_head_dim, cfg.n_groups, cfg.ssm_state_size, nondegenerate=False)
mixer = tv.Mamba2MixerPrefill(cfg).to(torch.bfloat16).cuda()
mixer.load_state_dict({k: v.to(torch.bfloat16) for k, v in p.items()})
proj = (torch.randn(1, L, cfg.projection_size, device="cuda") * 0.5).to(torch.bfloat16)
with torch.no_grad():
    for _ in range(5):
        tv.sharded_scan_core(mixer, proj)
    torch.cuda.synchronize(); dist.barrier()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            tv.sharded_scan_core(mixer, proj)
        torch.cuda.synchronize()
if rank <= 1:
    # device timeline of the last profiled step: start offset, duration, stream, kernel
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    per = len(evs) // 10
    last = evs[-per:]
    t0 = last[0].time_range.start
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/timeline_w{world}_rank{rank}.log", "w") as f:
        f.write(f"== rank {rank} of {world}, {L} tokens/rank: timeline of one step (us): start  dur  name\n")
        for e in last:
            f.write(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:70]}\n")
    if rank == 0: print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
dist.destroy_process_group()
