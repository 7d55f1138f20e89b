"""Per-kernel device times (CUDA events, warm) for the S1/S2 path at a benchmark shape."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=15_500_000)
ap.add_argument("--cols", type=int, default=833)
ap.add_argument("--states", type=int, default=18)
ap.add_argument("--kind", default="realistic")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()

x = synth.synth_states_device(a.bins, a.cols, a.states, seed=1, kind=a.kind)
cnt = torch.empty((a.bins, a.states), dtype=torch.int16, device="cuda")
out = torch.empty((a.bins, a.states), dtype=torch.float32, device="cuda")


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
res["k1_counts_ms"] = timeit(lambda: engine.bin_counts(x, a.cols, a.states, out=cnt))
n1, n2 = engine.expected_tables(cnt, a.cols)
res["k2_expected_s1s2_ms"] = timeit(lambda: engine.expected_tables(cnt, a.cols))
res["k2_expected_s2_only_ms"] = timeit(lambda: engine.expected_tables(cnt, a.cols, want_s1=False))
e1, e2 = engine.normalize(n1), engine.normalize(n2)
res["k4_normalize_ms"] = timeit(lambda: engine.normalize(n2))
res["k5_s1_table_ms"] = timeit(lambda: engine.scores_s1(cnt, a.cols, e1, out32=out))
res["k5_s2_table_ms"] = timeit(lambda: engine.scores_s2(cnt, a.cols, e2, out32=out))
sub = cnt[: a.bins // 8].contiguous()
sub_out = out[: a.bins // 8]
res["k5_s1_direct_ms_per_full"] = 8 * timeit(lambda: engine.scores_s1(sub, a.cols, e1, out32=sub_out, mode=1))
res["k5_s2_direct_ms_per_full"] = 8 * timeit(lambda: engine.scores_s2(sub, a.cols, e2, out32=sub_out, mode=1))
gb = a.bins * a.cols / 1e9
res["k1_GBps"] = gb / (res["k1_counts_ms"] * 1e-3)
print(json.dumps({"shape": vars(a), **res}))
