"""Bring-up check of the tcgen05 one-hot Gram kernel against a torch reference on the same GPU."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine  # noqa: E402
from oracle import epilogos_oracle as orc  # noqa: E402


def check(bins, cols, k, seed=0, budget=24 << 30):
    x = orc.synth_states(bins, cols, k, seed, kind="uniform")
    xd = engine.pack_states(x).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tiles, plan = engine.s3_expected_tiles(xd, cols, k, onehot_budget_bytes=budget)
    counts, exp = engine.s3_finalize(tiles, cols, k, plan["mp"], bins)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # torch reference: one-hot Gram in fp32 (exact below 2^24)
    oh = torch.zeros((bins, cols * k), dtype=torch.float32, device="cuda")
    idx = (torch.arange(cols, device="cuda")[None, :] * k + xd[:, :cols].long())
    oh.scatter_(1, idx, 1.0)
    g = (oh.T @ oh).round().long().reshape(cols, k, cols, k).permute(0, 2, 1, 3).contiguous()
    g[torch.arange(cols), torch.arange(cols)] = 0
    ok = torch.equal(g, counts)
    nbad = int((g != counts).sum())
    e_ref = (g.double() / float(bins * cols * (cols - 1))).float()
    ok_e = torch.equal(e_ref, exp)
    print("bins=%d cols=%d k=%d tiles=%d mp=%d  counts_equal=%s (bad=%d of %d) exp_equal=%s  %.1f ms" % (
        bins, cols, k, plan["ntiles"], plan["mp"], ok, nbad, g.numel(), ok_e, dt * 1e3), flush=True)
    if not ok:
        bad = (g != counts).nonzero()[:5]
        for b in bad:
            b = tuple(int(v) for v in b)
            print("   first mismatches", b, int(g[b]), int(counts[b]))
    return ok and ok_e


if __name__ == "__main__":
    allok = True
    allok &= check(128, 12, 15)
    allok &= check(300, 12, 15)
    allok &= check(1000, 40, 18)
    allok &= check(5000, 100, 18)
    allok &= check(3000, 100, 18, budget=100 * 18 * 1024)      # forces bin chunks + accumulate
    allok &= check(2048, 833, 18)
    print("ALL OK" if allok else "FAILED")
