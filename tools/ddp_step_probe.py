"""Where does the multi-GPU step lose time?  Variants of the bench step under torchrun."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
bins, cols, k = 15_500_000, 833, 18
x = synth.synth_states_device(bins, cols, k, seed=1234 + rank)
cnt = torch.empty((bins, k), dtype=torch.int16, device="cuda")
scores = torch.empty((bins, k), dtype=torch.float32, device="cuda")


def run(name, allreduce=True, sleep_before=0.0, barrier_each=False):
    def step():
        engine.bin_counts(x, cols, k, out=cnt)
        _, n2 = engine.expected_tables(cnt, cols, want_s1=False)
        if allreduce:
            dist.all_reduce(n2)
        e = engine.normalize(n2)
        engine.scores_s2(cnt, cols, e, out32=scores)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        step()
    t_cpu = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-28s %.3f ms/step (cpu enqueue %.3f ms/step)" % (name, ms.item(), t_cpu / 20 * 1e3), flush=True)


run("with all_reduce")
run("without all_reduce", allreduce=False)
run("with all_reduce (again)")
os.environ["NCCL_PROTO"] = "LL"
run("with all_reduce (3rd)")
dist.destroy_process_group()
