"""Multi-GPU correctness of the stage drivers (run under torchrun with >= 2 GPUs):
expected -> expectedCombination -> scores for S1/S2/S3 and paired S1 with rows sharded over the ranks must give
the same files as the committed single-process reference goldens."""
import gzip
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from epilogos_b200 import expected, expectedCombination, scores, session  # noqa: E402
from test_host_stages import write_tsv  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
g = np.load(ROOT / "tests" / "golden" / "real10_chr1_k18.npz")
gp = np.load(ROOT / "tests" / "golden" / "paired_real10_k18.npz")
obj = [tempfile.mkdtemp() if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
tmp = Path(obj[0])
ok = True


def close32(a, b):
    neq = a != b
    if not neq.any():
        return True
    ulp = np.spacing(np.maximum(np.abs(a[neq]), np.abs(b[neq])))
    return bool(np.all(np.abs(a[neq].astype(np.float64) - b[neq]) <= ulp * 1.0000001) and neq.mean() < 2e-3)


for s in (1, 2, 3):
    session.clear()
    out = tmp / ("out%d" % s)
    f = tmp / ("in%d" % s) / "epilogos_matrix_chr1.txt.gz"
    if rank == 0:
        out.mkdir(); f.parent.mkdir()
        write_tsv(f, g["x"], gz=True)
    dist.barrier()
    tag = "in_s%d" % s
    expected.main(f, "null", 18, s, out, tag, 1, False)
    if rank == 0:
        counts = np.load(out / ("temp_exp_freq_%s_epilogos_matrix_chr1.npy" % tag))
        ok &= bool(np.array_equal(counts, g["s%d_counts" % s]))
    expectedCombination.main(out, out / ("exp_freq_%s.npy" % tag), tag, False)
    scores.main(f, "null", 18, s, out, out / ("exp_freq_%s.npy" % tag), tag, 1, 17, -1, False)
    if rank == 0:
        exp = np.load(out / ("exp_freq_%s.npy" % tag))
        npz = np.load(out / ("temp_scores_%s_epilogos_matrix_chr1.npz" % tag), allow_pickle=True)
        e_ok = exp.tobytes() == g["s%d_exp" % s].tobytes()
        sc_ok = close32(npz["scoreArr"], g["s%d_scores" % s]) if s < 3 else \
            float(np.max(np.abs(npz["scoreArr"] - g["s3_scores"]))) < 2e-2
        print("S%d world=%d counts+exp bit-exact=%s scores=%s" % (s, world, ok and e_ok, sc_ok), flush=True)
        ok &= e_ok and sc_ok
# paired S1, reference-style null replay is single-process only; check delta + quiescence + expected here
session.clear()
out = tmp / "outp"
fa, fb = tmp / "a" / "epilogos_matrix_chr1.txt", tmp / "b" / "epilogos_matrix_chr1.txt"
if rank == 0:
    out.mkdir(); fa.parent.mkdir(); fb.parent.mkdir()
    write_tsv(fa, gp["xa"]); write_tsv(fb, gp["xb"])
dist.barrier()
expected.main(fa, fb, 18, 1, out, "a_b_s1", 1, False)
expectedCombination.main(out, out / "exp_freq_a_b_s1.npy", "a_b_s1", False)
scores.main(fa, fb, 18, 1, out, out / "exp_freq_a_b_s1.npy", "a_b_s1", 1, 17, -1, False)
if rank == 0:
    exp = np.load(out / "exp_freq_a_b_s1.npy")
    q = np.load(out / "temp_quiescence_a_b_s1_epilogos_matrix_chr1.npz")["quiescenceArr"]
    nd = np.load(out / "temp_nullDistances_a_b_s1_epilogos_matrix_chr1.npz")["nullDistances"]
    with gzip.open(out / "pairwiseDelta_a_b_s1_epilogos_matrix_chr1.txt.gz", "rb") as z:
        lines = z.read().split(b"\n")
    ref_lines = gp["s1_delta_text"].tobytes().split(b"\n")
    p_ok = exp.tobytes() == gp["s1_exp"].tobytes() and np.array_equal(q, gp["s1_quiescence"]) and \
        len(lines) == len(ref_lines) and sum(a != b for a, b in zip(lines, ref_lines)) <= 3 and nd.shape == gp["s1_null"].shape
    print("paired S1 world=%d exp+quiescence+delta text ok=%s" % (world, p_ok), flush=True)
    ok &= bool(p_ok)
    print("MGPU ALL OK" if ok else "MGPU FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
