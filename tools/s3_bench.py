"""Device timing of the S3 expected path (one-hot expansion, tcgen05 Gram, finalise) at a benchmark shape."""
import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=1_250_000)
ap.add_argument("--cols", type=int, default=833)
ap.add_argument("--states", type=int, default=18)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--chunk", type=int, default=131072)
ap.add_argument("--score-bins", type=int, default=65536)
a = ap.parse_args()

x = synth.synth_states_device(a.bins, a.cols, a.states, seed=3)
plan = engine.s3_plan(a.bins, a.cols, a.states)
mp, bp = plan["mp"], plan["bp"]
oht = torch.empty(mp * bp, dtype=torch.int8, device="cuda")
tiles = torch.empty(plan["tile_bytes"] // 4, dtype=torch.int32, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())


def ev(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def chunked():
    engine.s3_expected_tiles(x, a.cols, a.states, tiles=tiles, chunk_bins=a.chunk)


t_chunked = ev(chunked)
cb = min(a.chunk, bp) // 128 * 128
t_probe = ev(lambda: _lib.call("epi_s3_gram", p(oht), mp, cb, p(tiles), 2, st))
t_oh = ev(lambda: _lib.call("epi_s3_onehot", p(x), a.bins, a.cols, x.shape[1], a.states, p(oht), mp, bp, st))
t_gram = ev(lambda: _lib.call("epi_s3_gram", p(oht), mp, bp, p(tiles), 0, st))
t_fin = ev(lambda: engine.s3_finalize(tiles, a.cols, a.states, mp, a.bins, want_counts=False, want_exp=True))
_, exp3 = engine.s3_finalize(tiles, a.cols, a.states, mp, a.bins, want_counts=False, want_exp=True)
t_terms = ev(lambda: engine.s3_terms(exp3.reshape(-1), a.cols, a.states))
terms = engine.s3_terms(exp3.reshape(-1), a.cols, a.states)
xs = x[: a.score_bins]
so = torch.empty((xs.shape[0], a.states), dtype=torch.float32, device="cuda")
t_score = ev(lambda: engine.scores_s3(xs, a.cols, a.states, terms, out32=so))
ck = a.cols * a.states
useful = a.bins * ck * (ck + 1)                      # upper triangle incl. diagonal, 2 ops per MAC
issued = plan["ntiles"] * 128 * 256 * 2 * bp          # what the tensor cores actually execute
print(json.dumps({"shape": vars(a), "plan": plan, "onehot_ms": t_oh, "gram_ms": t_gram, "finalize_ms": t_fin,
                  "useful_TOPS": useful / (t_gram * 1e-3) / 1e12, "issued_TOPS": issued / (t_gram * 1e-3) / 1e12,
                  "chunked_expected_ms": t_chunked, "chunked_useful_TOPS": useful / (t_chunked * 1e-3) / 1e12,
                  "probe_ms": t_probe, "probe_issued_TOPS": plan["ntiles"] * 128 * 256 * 2 * cb / (t_probe * 1e-3) / 1e12,
                  "terms_ms": t_terms, "score_ms_for_score_bins": t_score, "score_bins_per_s": xs.shape[0] / (t_score * 1e-3),
                  "score_lookups_per_s": xs.shape[0] * a.cols * (a.cols - 1) / (t_score * 1e-3), "onehot_GBps": mp * bp / (t_oh * 1e-3) / 1e9,
                  "bins_per_s_expected": a.bins / ((t_oh + t_gram + t_fin) * 1e-3)}))
