"""Fixed cost of the S1 / S2 step: the CUDA-graph-captured step (K1, K2, K4, K5; no all-reduce) on one GPU at 1/8, 1/4,
1/2 and 1/1 of a genome -- the per-GPU shares of the strong-scaling target at 8, 4, 2 and 1 GPUs."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402

cols, k = 833, 18
full = 15_500_000
x = synth.synth_states_device(full, cols, k, seed=1)
res = {}
for sal in (2, 1):
    for div in (8, 4, 2, 1):
        rows = full // div
        xs = x[:rows]
        cnt = torch.empty((rows, k), dtype=torch.int16, device="cuda")
        out = torch.empty((rows, k), dtype=torch.float32, device="cuda")

        def step():
            engine.bin_counts(xs, cols, k, out=cnt)
            n1, n2 = engine.expected_tables(cnt, cols, want_s1=sal == 1, want_s2=sal == 2)
            e = engine.normalize(n1 if sal == 1 else n2)
            (engine.scores_s1 if sal == 1 else engine.scores_s2)(cnt, cols, e, out32=out)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res["s%d_1/%d" % (sal, div)] = round(e0.elapsed_time(e1) / 50, 4)
        del g
print(json.dumps(res))
