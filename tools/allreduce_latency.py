"""Latency of the path's exchange step (all-reduce of the int64 expected table) under torchrun."""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(324, dtype=torch.int64, device="cuda")
for n in (324, 324 * 1024):
    t = torch.ones(n, dtype=torch.int64, device="cuda")
    for _ in range(20):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        print("all_reduce int64[%d]: %.1f us per call" % (n, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
# with a 3 ms kernel in between (as in the bench step)
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
t = torch.ones(324, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    x.mul_(1.0001)
e1.record()
torch.cuda.synchronize()
base = e0.elapsed_time(e1) / 50
e0.record()
for _ in range(50):
    x.mul_(1.0001)
    dist.all_reduce(t)
e1.record()
torch.cuda.synchronize()
if dist.get_rank() == 0:
    print("kernel alone %.3f ms, kernel + all_reduce %.3f ms" % (base, e0.elapsed_time(e1) / 50), flush=True)
dist.destroy_process_group()
