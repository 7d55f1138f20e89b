"""Timing of the similarity-search distance engine (f4) at a genome-like shape: reduced genome of G bins x K states,
R regions of interest of nS reduced bins, nDesired matches.  Reports per-stage device times and ROIs per second, and
(with --cpu) the oracle's time for one ROI on the host (= the reference's algorithm: sklearn-style distances, mode, argsort)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import similaritySearch_calc as ssc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=3_100_000)
ap.add_argument("--states", type=int, default=18)
ap.add_argument("--ns", type=int, default=5)
ap.add_argument("--rois", type=int, default=64)
ap.add_argument("--desired", type=int, default=100)
ap.add_argument("--cpu", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(0)
red = rng.random((a.genome, a.states)) * rng.random((a.genome, 1))
red[rng.random(a.genome) < 0.5] = 0.01
starts = rng.integers(0, a.genome - a.ns, a.rois)
cube = np.stack([red[s:s + a.ns] for s in starts])
genome = torch.from_numpy(red).cuda()
xx = ssc.row_norms(genome)


def timed(fn, reps=3):
    out = fn()                                            # warm-up (allocations)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    return out, best


rois = torch.from_numpy(cube[:ssc.ROI_BATCH]).cuda()
ssc.window_distances(genome, xx, rois)
dist, t_dist = timed(lambda: ssc.window_distances(genome, xx, rois))
(svals, sidx), t_sort = timed(lambda: torch.sort(dist, dim=1, stable=True))
mode, t_mode = timed(lambda: ssc.mode_of_sorted(svals))
res = {"shape": vars(a), "batch": ssc.ROI_BATCH, "distances_ms_per_batch": t_dist, "sort_ms_per_batch": t_sort,
       "mode_ms_per_batch": t_mode,
       "distance_GFLOPs": 2.0 * a.states * a.ns * (a.genome - a.ns + 1) * ssc.ROI_BATCH / (t_dist * 1e-3) / 1e9}
starts_dev = torch.from_numpy(starts.astype(np.int64)).cuda()
picks, t_pick = timed(lambda: ssc.pick(svals, sidx, mode, starts_dev[:ssc.ROI_BATCH], a.ns, a.desired))
res["pick_ms_per_batch"] = t_pick
torch.cuda.synchronize()
t = time.time()
n = 0
for b0 in range(0, a.rois, ssc.ROI_BATCH):
    r = torch.from_numpy(cube[b0:b0 + ssc.ROI_BATCH]).cuda()
    d = ssc.window_distances(genome, xx, r)
    sv, si = torch.sort(d, dim=1, stable=True)
    m = ssc.mode_of_sorted(sv)
    ssc.pick(sv, si, m, starts_dev[b0:b0 + r.shape[0]], a.ns, a.desired).cpu()
    n += r.shape[0]
res["rois_per_s_end_to_end"] = n / (time.time() - t)
if a.cpu:
    from oracle import simsearch_oracle as so
    t = time.time()
    so.similar_regions(red, cube[0], int(starts[0]), a.desired)
    res["cpu_oracle_s_per_roi"] = time.time() - t
print(json.dumps(res))
