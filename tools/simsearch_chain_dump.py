"""Run the simsearch build chain on the real-data fixture on the GPU and dump the index array (tie analysis)."""
import sys, numpy as np, tempfile
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_simsearch_prep import _write_scores
from epilogos_b200 import similaritySearch_run as ssr
g = np.load("tests/golden/simsearch_prep_real_chr1_60k.npz")
with tempfile.TemporaryDirectory() as tmp:
    tmp = Path(tmp)
    _write_scores(tmp / "scores_x.txt.gz", g)
    out = tmp / "build"; out.mkdir()
    idx = ssr.buildSimSearch(tmp / "scores_x.txt.gz", out, -1, 100, -1, -1.0)
    np.save("gpurun_out/chain_idx.npy", idx)
c = np.load("tests/golden/simsearch_chain_real_chr1_60k.npz")
bad = np.flatnonzero((idx != c["indices"]).any(axis=1))
print("regions differing:", len(bad), "of", len(idx))
for r in bad[:5]:
    k = np.flatnonzero(idx[r] != c["indices"][r])
    print(r, "first diff at", k[0], "ours", idx[r][k[0]:k[0]+6], "ref", c["indices"][r][k[0]:k[0]+6], "same set", set(idx[r]) == set(c["indices"][r]))
