"""`simsearch -b` under torchrun: the regions are split over the ranks (one GPU each) the way the reference splits them
over SLURM jobs; rank 0 prepares and writes.  Compares the combined index array with a single-GPU result.

    torchrun --nproc-per-node 2 tools/simsearch_mgpu_check.py SCORES.txt.gz SINGLE_GPU_INDICES.npy
"""
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from epilogos_b200 import dist, similaritySearch_run as ssr
    dist.init_from_env()
    out = Path(tempfile.gettempdir()) / "simsearch_mgpu_check"
    if dist.rank() == 0:
        shutil.rmtree(out, ignore_errors=True)
        out.mkdir(parents=True)
    dist.barrier()
    t = time.time()
    idx = ssr.buildSimSearch(sys.argv[1], out, -1, 100, -1, -1.0)
    if dist.rank() == 0:
        want = np.load(sys.argv[2])
        same = idx.shape == want.shape and np.array_equal(idx, want)
        print("SIMSEARCH MGPU world=%d: %d regions in %.2f s, equal to the single-GPU indices: %s"
              % (dist.world_size(), len(idx), time.time() - t, same), flush=True)
        left = sorted(p.name for p in out.iterdir())
        print("files:", left, flush=True)
        if not same:
            sys.exit(1)
    dist.barrier()


if __name__ == "__main__":
    main()
