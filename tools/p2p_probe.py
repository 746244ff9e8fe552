"""torchrun helper (2+ GPUs): how fast can one rank read / write 5.24 MB of a peer's symmetric memory?
pull = local.copy_(peer), push = peer.copy_(local), fold = the fused exchange+fold kernel reading rank-1's summary."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
import timeviper_b200 as tv

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 128 * 80 * 128
buf = symm.empty(n + 128, dtype=torch.float32, device="cuda")
hdl = symm.rendezvous(buf, dist.group.WORLD)
buf.normal_()
loc = torch.empty_like(buf)
peer = hdl.get_buffer((rank + 1) % world, (n + 128,), torch.float32, 0)
hdl.barrier(channel=0)


def timeit(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


res = {}
for name, fn in (("pull copy_", lambda: loc.copy_(peer)), ("push copy_", lambda: peer.copy_(loc)),
                 ("local copy_", lambda: loc.copy_(buf))):
    dist.barrier()
    if rank == 0:
        res[name] = timeit(fn)
    dist.barrier()
ptrs = [int(p) for p in hdl.buffer_ptrs]
sp = [ptrs[(rank + 1) % world]] * 2
lp = [p + n * 4 for p in sp]
for r in (1, 2):
    dist.barrier()
    if rank == 0:
        res[f"fold_p2p rank={r} (reads {r} x 5.24 MB from the peer)"] = timeit(
            lambda: tv.ops.fold_boundary_states_p2p(sp, lp, r, (1, 128, 80, 128), torch.device("cuda", local)))
    dist.barrier()
if rank == 0:
    for k, v in res.items():
        print(f"{k:60s} {v:8.1f} us   {5.24e6 / v / 1e3:7.1f} GB/s per 5.24 MB")
dist.destroy_process_group()
