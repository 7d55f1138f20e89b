"""Small shapes through every kernel of the S1/S2/S3 path (for compute-sanitizer memcheck runs)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine  # noqa: E402

rng = np.random.default_rng(0)
for bins, cols, k in ((300, 40, 18), (129, 300, 15), (70, 13, 6)):
    x = rng.integers(0, k, size=(bins, cols)).astype(np.int8)
    xd = engine.pack_states(x).cuda()
    cnt = engine.bin_counts(xd, cols, k)
    n1, n2 = engine.expected_tables(cnt, cols)
    e1, e2 = engine.normalize(n1), engine.normalize(n2 + 1)
    engine.scores_s1(cnt, cols, e1, want64=True)
    engine.scores_s2(cnt, cols, e2, want64=True)
    engine.scores_s2(cnt, cols, engine.normalize(n2), want64=True, mode=1)
    if cols <= 40:
        tiles, plan = engine.s3_expected_tiles(xd, cols, k)
        counts, exp = engine.s3_finalize(tiles, cols, k, plan["mp"], bins)
        terms = engine.s3_terms(exp.reshape(-1), cols, k)
        engine.scores_s3(xd, cols, k, terms, want64=True)
    torch.cuda.synchronize()
    print("ok", bins, cols, k, flush=True)
print("done")
