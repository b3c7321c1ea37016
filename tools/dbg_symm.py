"""Feasibility probe: torch symmetric memory (peer pointers + barrier) on this box."""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
import torch.distributed._symmetric_memory as symm
t = symm.empty(1024, dtype=torch.int64, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
t.fill_(rank + 1)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.int64)
print(rank, "peer value", int(peer[0]), flush=True)
# latency of barrier and of a tiny NCCL all_reduce / all_gather
x = torch.zeros(8, dtype=torch.int64, device=dev)
g = torch.zeros(8 * world, dtype=torch.int64, device=dev)
for name, fn in (("symm barrier", lambda: hdl.barrier()), ("nccl all_reduce 64B", lambda: dist.all_reduce(x)),
                 ("nccl all_gather 64B", lambda: dist.all_gather_into_tensor(g, x))):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(name, a.elapsed_time(b) / 200 * 1000, "us", flush=True)
dist.barrier(); dist.destroy_process_group()
