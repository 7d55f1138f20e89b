"""File-to-file wall clock of the drop-in CLI at REAL width: `epilogos -l -i DIR -j STATES -o OUT -s S` on a synthetic
BINS x 833 x 18 matrix written as TSV.gz (README.md:286-292), with the host-side stage breakdown (EPILOGOS_B200_TIMING):
inflate + parse, host -> device, kernels, device -> host, %.5f formatting + deflate, npz hand-over files.  Optionally the
unmodified reference (oracle/_ref) on a row subset of the same file for the CPU comparison.

    python tools/cli_width833.py [--bins 400000] [--files 1] [--saliency 1 2] [--reference-bins 20000]
    torchrun --nproc-per-node N tools/cli_width833.py ...     (rows mode; EPILOGOS_B200_SHARD=files with --files >= N)
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["EPILOGOS_B200_TIMING"] = "1"

from click.testing import CliRunner  # noqa: E402

from epilogos_b200 import dist, run, session, timing  # noqa: E402
from oracle import epilogos_oracle as orc  # noqa: E402   (input generation only)
from oracle import reference_driver as ref  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=400_000)
ap.add_argument("--cols", type=int, default=833)
ap.add_argument("--states", type=int, default=18)
ap.add_argument("--files", type=int, default=1)
ap.add_argument("--saliency", type=int, nargs="+", default=[1, 2])
ap.add_argument("--reference-bins", type=int, default=0)
ap.add_argument("--dir", default="")
a = ap.parse_args()

META = "zero_index\tone_index\tshort_name\tlong_name\n" + "".join(
    "%d\t%d\tS%d\tstate %d\n" % (i, i + 1, i + 1, i + 1) for i in range(a.states))
# Input generation happens BEFORE the process group exists: rank 0 writes the files, the others wait for a sentinel file.
# (Joining the group first and letting rank 0 arrive a minute late at the first collective made the NCCL communicator
# bootstrap of the other ranks time out on an 8-GPU box.)
env_rank = int(os.environ.get("RANK", "0"))
base = Path(a.dir) if a.dir else Path(tempfile.gettempdir()) / "epi_cli_width"
inp = base / "in"
ready = base / ("ready_%d_%d_%d" % (a.bins, a.files, a.cols))
res = {"bins": a.bins, "biosamples": a.cols, "states": a.states, "files": a.files,
       "world": int(os.environ.get("WORLD_SIZE", "1")), "shard": os.environ.get("EPILOGOS_B200_SHARD", "rows")}
if env_rank == 0:
    inp.mkdir(parents=True, exist_ok=True)
    t = time.time()
    per = a.bins // a.files
    for f in range(a.files):
        path = inp / ("epilogos_matrix_chr%d.txt.gz" % (f + 1))
        if not path.exists():
            ref.write_matrix_tsv_gz(path, orc.synth_states(per, a.cols, a.states, seed=50 + f), chrom="chr%d" % (f + 1))
    (base / "meta.tsv").write_text(META)
    res["write_input_s"] = round(time.time() - t, 1)
    res["input_MB"] = round(sum(p.stat().st_size for p in inp.glob("*")) / 1e6, 1)
    ready.write_text("ok")
else:
    t = time.time()
    while not ready.exists():
        time.sleep(0.2)
        if time.time() - t > 900:
            raise SystemExit("input files never appeared")
rank, world = env_rank, int(os.environ.get("WORLD_SIZE", "1"))
# run.main joins the process group of a torchrun launch itself and destroys it when it returns: under torchrun the CLI
# is invoked exactly ONCE per process (one saliency, cold), as a user would; a plain `python` launch repeats it warm.
multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
if multi:
    a.saliency = a.saliency[:1]
for s in a.saliency:
    for rep in range(1 if multi else 2):         # second run: library, CUDA context and page cache warm
        session.clear()
        timing.reset()
        out = base / ("out_s%d_%d_%s" % (s, rep, res["shard"]))
        t = time.time()
        r = CliRunner().invoke(run.main, ["-l", "-i", str(inp), "-o", str(out), "-j", str(base / "meta.tsv"), "-s", str(s)])
        dt = time.time() - t                     # run.main ends with a barrier: every rank sees the whole job's wall time
        assert r.exit_code == 0, r.output + repr(r.exception)
    res["s%d_wall_s" % s] = round(dt, 2)
    res["s%d_bins_per_s" % s] = round(a.bins / dt)
    res["s%d_rank0_stages_s" % s] = timing.report()
if rank == 0 and a.reference_bins:
    # the unmodified reference on the first rows of the same kind of file, all host cores (what `epilogos -l -c 0` runs)
    import gzip
    sub = base / "ref_in"
    sub.mkdir(exist_ok=True)
    src = sorted(inp.glob("*"))[0]
    with gzip.open(src, "rb") as g, gzip.open(sub / "epilogos_matrix_chr1.txt.gz", "wb", compresslevel=1) as o:
        for i, line in enumerate(g):
            if i >= a.reference_bins:
                break
            o.write(line)
    for s in a.saliency:
        wall, _ = ref.time_pipeline(sub / "epilogos_matrix_chr1.txt.gz", "null", a.states, s, os.cpu_count(), base)
        res["reference_s%d_bins_per_s" % s] = round(a.reference_bins / wall)
    res["reference_cores"] = os.cpu_count()
if rank == 0:
    print(json.dumps(res))
