"""Per-process residency of parsed inputs and per-bin counts.

The reference re-parses every input file in every stage and in every worker (helpers.py:152-155).  Here the
rank's row range of a file is parsed once, packed to int8, counted once on the GPU, and the device-resident
counts (2K bytes per bin) are reused by the score stage of the same process.
"""
from pathlib import Path

import numpy as np

from . import dist, helpers

_backend = None
_cache = {}

# In-process hand-over of the score stage to the region-of-interest stage.  The reference passes the scores of every file
# through temp_scores_<tag>_<file>.npz (scores.py:166-169), which the ROI stage reads and deletes (roiSingle.py:43-76).  When
# both stages run in THIS process (the CLI, run.run_stages) the arrays are handed over in memory instead: the file would
# be written, read back and removed within seconds (0.3-0.6 s of a 2 s run at 400 k bins x 833).  Stand-alone calls of
# scores.main still write the file, and EPILOGOS_B200_KEEP_TEMP=1 forces it everywhere.
handover = {}              # str(path of the temp_scores npz that was not written) -> dict(chrName, scoreArr, chrom, start, end)
handover_enabled = False


def get_backend(backend=None):
    global _backend
    if backend is not None:
        return backend
    if _backend is None:
        from .backend import CudaBackend
        _backend = CudaBackend()
    return _backend


class Shard:
    """The rows [lo, hi) of one input file (or file pair) owned by this rank."""

    def __init__(self, file1, file2, num_states, backend, rank=0, world=0):
        self.backend = get_backend(backend)
        self.num_states = num_states
        if not Path(file1).is_file():
            raise FileNotFoundError(str(file1))
        # one pass over the file: the row count comes out of the parse, this rank's range is cut from the parsed rows
        pinned = getattr(self.backend, "name", "") == "cuda"
        split = (rank, world) if world else (dist.rank(), dist.world_size())
        self.loc, self.states_a, self.total_rows = helpers.read_matrix(file1, None, want_locations=True,
                                                                       num_states=num_states, pinned=pinned,
                                                                       split=split, return_total=True)
        self.lo, self.hi = helpers.splitRows(self.total_rows, split[1])[split[0]]
        rows = (self.lo, self.hi)
        self.states_b = None
        if str(file2) != "null":
            _, self.states_b = helpers.read_matrix(file2, rows, want_locations=False, num_states=num_states)
            if self.states_b.shape[0] != self.states_a.shape[0]:
                raise ValueError("paired input files must have the same number of rows")
        self._counts = {}
        self._states_dev = None

    @property
    def paired(self):
        return self.states_b is not None

    @property
    def width(self):
        """Biosamples of the matrix the expected table is computed over: the union of both groups in paired
        mode (helpers.py:173-179)."""
        return self.states_a.shape[1] + (self.states_b.shape[1] if self.paired else 0)

    def combined(self):
        return self.states_a if not self.paired else np.concatenate((self.states_a, self.states_b), axis=1)

    def counts(self, which="all"):
        """Per-bin counts of the whole matrix ("all": the union [A | B] in paired mode) or of one group."""
        if which not in self._counts:
            if which == "all" and self.paired:
                self._counts[which] = self.backend.add_counts(self.counts("a"), self.counts("b"))
            else:
                m = self.states_b if which == "b" else self.states_a
                self._counts[which] = self.backend.counts(m, self.num_states)
        return self._counts[which]

    def states_device(self):
        if self._states_dev is None:
            self._states_dev = self.backend.states_to_device(self.combined())
        return self._states_dev


def _key(file1, file2):
    p1 = Path(file1).resolve()
    st = p1.stat()
    return (str(p1), st.st_mtime_ns, st.st_size, str(file2), dist.rank(), dist.world_size())


_keep = 4


def load_shard(file1, file2, num_states, backend=None, keep=None):
    key = _key(file1, file2)
    if key not in _cache:
        while len(_cache) >= (keep or _keep):
            _cache.pop(next(iter(_cache)))
        _cache[key] = Shard(file1, file2, num_states, backend)
    return _cache[key]


def prefetch(pairs, num_states, backend=None, workers=None):
    """Parse this rank's rows of ALL input files concurrently, before the stages run.  The reference parses every file
    again in every stage and worker; the native packer (csrc/hostio.cu) releases the GIL and gives every file its share of
    the cores (inflate | parse threads), so a whole-genome directory of per-chromosome files is read in the time of its largest files
    instead of their sum.  The cache is sized to hold every file, so the score stage re-uses the parsed matrices and the
    device-resident counts of the expected stage."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    global _keep
    pairs = list(pairs)
    _keep = max(_keep, len(pairs) + 1)
    todo = [(f, f2) for f, f2 in pairs if _key(f, f2) not in _cache]
    if not todo:
        return
    be = get_backend(backend)
    rank, world = dist.rank(), dist.world_size()          # resolved on the calling thread
    workers = workers or max(1, min(len(todo), (os.cpu_count() or 2) // 2))
    from . import _lib
    _lib.call("epi_reader_concurrency", int(workers))     # every reader takes 1/workers of this rank's cores from the start
    try:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            shards = list(pool.map(lambda p: Shard(p[0], p[1], num_states, be, rank=rank, world=world), todo))
    finally:
        _lib.call("epi_reader_concurrency", 0)
    for (f, f2), sh in zip(todo, shards):
        _cache[_key(f, f2)] = sh


def clear():
    _cache.clear()
