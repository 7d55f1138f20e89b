"""Device timing of paired mode (BASELINE configs[3]: two groups, S1, P null shuffles per bin) on one GPU."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=1_250_000)
ap.add_argument("--c1", type=int, default=400)
ap.add_argument("--c2", type=int, default=433)
ap.add_argument("--states", type=int, default=18)
ap.add_argument("--saliency", type=int, default=1)
ap.add_argument("--perms", type=int, default=1000)
ap.add_argument("--batch", type=int, default=20)
a = ap.parse_args()
k = a.states
xa = synth.synth_states_device(a.bins, a.c1, k, seed=5)
xb = synth.synth_states_device(a.bins, a.c2, k, seed=6)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def real():
    ca = engine.bin_counts(xa, a.c1, k)
    cb = engine.bin_counts(xb, a.c2, k)
    comb = (ca + cb).contiguous()
    n1, n2 = engine.expected_tables(comb, a.c1 + a.c2, want_s1=a.saliency == 1, want_s2=a.saliency == 2)
    e = engine.normalize(n1 if a.saliency == 1 else n2)
    score = (lambda c, w: engine.scores_s1(c, w, e)) if a.saliency == 1 else \
        (lambda c, w: engine.scores_s2(c, w, e))
    sa, sb = score(ca, a.c1), score(cb, a.c2)
    delta, _ = engine.pairwise_combine(sa, sb, None, None)
    quies = engine.quiescent_mask(ca, a.c1, cb, a.c2, k - 1)
    dist, md = engine.pairwise_real_reduce(delta)
    return ca, cb, e, score, delta, quies, dist


real()
(ca, cb, e, score, delta, quies, dist), t_real = timed(real)


def null(nperm, seed):
    oa, ob = engine.shuffled_counts_philox(ca, cb, a.c1, a.c2, seed, nperm)
    na = score(oa.reshape(-1, k), a.c1)
    nb = score(ob.reshape(-1, k), a.c2)
    _, nd = engine.pairwise_combine(None, None, na, nb)
    return nd.reshape(nperm, a.bins)


null(1, 0)
_, t_null1 = timed(lambda: null(1, 1))
done, t_nullp = 0, 0.0
while done < a.perms:
    n = min(a.batch, a.perms - done)
    _, t = timed(lambda: null(n, 100 + done))
    t_nullp += t
    done += n
print(json.dumps({"shape": vars(a), "real_ms": t_real, "null_1perm_ms": t_null1, "null_%dperm_ms" % a.perms: t_nullp,
                  "bins_per_s_real_plus_1perm": a.bins / ((t_real + t_null1) * 1e-3),
                  "bin_perms_per_s": a.bins * a.perms / (t_nullp * 1e-3)}))
