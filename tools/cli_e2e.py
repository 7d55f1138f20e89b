"""Wall-clock of the drop-in CLI on the real chr1 matrix (1,246,253 bins x 10 biosamples, 18 states; the stand-in for
BASELINE configs[0]): `epilogos -l -i DIR -j STATES -o OUT -s S` = gz/TSV parse -> expected -> combine -> scores -> gzip text
-> regions of interest, all in one process on one GPU.  The input is rebuilt from tests/golden/real10_chr1_full.npz
(the reference tree does not exist on the GPU box).  Compare with the unmodified reference's own wall time for the same
stages, measured in the authoring container (profiles/README.md)."""
import gzip
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from click.testing import CliRunner  # noqa: E402

from epilogos_b200 import run, session  # noqa: E402
from test_host_stages import write_tsv  # noqa: E402

META = "zero_index\tone_index\tshort_name\tlong_name\n" + "".join("%d\t%d\tS%d\tstate %d\n" % (i, i + 1, i + 1, i + 1)
                                                                      for i in range(18))
g = np.load(ROOT / "tests" / "golden" / "real10_chr1_full.npz")
x = g["x"]
res = {"bins": int(x.shape[0]), "biosamples": int(x.shape[1])}
with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    inp = d / "real10"
    inp.mkdir()
    t = time.time()
    write_tsv(inp / "epilogos_matrix_chr1.txt.gz", x, gz=True)
    res["write_input_s"] = round(time.time() - t, 2)
    meta = d / "meta.tsv"
    meta.write_text(META)
    for s in (1, 2):
        for rep in range(2):                 # second run: library, CUDA context and page cache warm
            session.clear()
            out = d / ("out_s%d_%d" % (s, rep))
            t = time.time()
            r = CliRunner().invoke(run.main, ["-l", "-i", str(inp), "-o", str(out), "-j", str(meta), "-s", str(s)])
            dt = time.time() - t
            assert r.exit_code == 0, r.output + repr(r.exception)
            res["cli_s%d_run%d_s" % (s, rep)] = round(dt, 2)
        with gzip.open(out / ("scores_real10_s%d_epilogos_matrix_chr1.txt.gz" % s), "rb") as f:
            n = sum(1 for _ in f)
        res["score_lines_s%d" % s] = n
        res["roi_lines_s%d" % s] = len((out / ("regionsOfInterest_real10_s%d.txt" % s)).read_text().splitlines())
print(json.dumps(res))
