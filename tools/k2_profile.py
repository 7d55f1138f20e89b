"""One launch each of K2 (tensor-core Gram tables) and the S2 score kernel at a shape (for ncu)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epilogos_b200 import engine, synth  # noqa: E402

bins = int(sys.argv[1]) if len(sys.argv) > 1 else 15_500_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 127
k = int(sys.argv[3]) if len(sys.argv) > 3 else 15
x = synth.synth_states_device(bins, cols, k, seed=1)
cnt = engine.bin_counts(x, cols, k)
out = torch.empty((bins, k), dtype=torch.float32, device="cuda")
for _ in range(2):
    engine.bin_counts(x, cols, k, out=cnt)
    n1, n2 = engine.expected_tables(cnt, cols, want_s1=False)
    e2 = engine.normalize(n2)
    engine.scores_s2(cnt, cols, e2, out32=out)
torch.cuda.synchronize()
print("done")
