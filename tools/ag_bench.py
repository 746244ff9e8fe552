"""torchrun helper: latency of the boundary-state all-gather (5.24 MB + 512 B per rank) under the NCCL settings in the
environment.  Rank 0 prints one line."""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                        pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
n = 128 * 80 * 128 + 128
send = torch.randn(n, device="cuda")
recv = torch.empty(world * n, device="cuda")
for _ in range(10):
    dist.all_gather_into_tensor(recv, send)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    dist.all_gather_into_tensor(recv, send)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 100 * 1e3], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NCCL_"))
    print(f"all-gather of {n * 4 / 1e6:.2f} MB x {world} ranks: {t.item():.1f} us   [{tag or 'defaults'}]", flush=True)
dist.destroy_process_group()
