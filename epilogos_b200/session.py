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

    def __init__(self, file1, file2, num_states, backend):
        self.backend = get_backend(backend)
        self.num_states = num_states
        self.total_rows = helpers.countRows(file1)
        self.lo, self.hi = helpers.splitRows(self.total_rows, dist.world_size())[dist.rank()]
        rows = (self.lo, self.hi)
        pinned = getattr(self.backend, "name", "") == "cuda"
        self.loc, self.states_a = helpers.read_matrix(file1, rows, want_locations=True, num_states=num_states,
                                                      pinned=pinned)
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


def load_shard(file1, file2, num_states, backend=None, keep=4):
    key = _key(file1, file2)
    if key not in _cache:
        while len(_cache) >= keep:
            _cache.pop(next(iter(_cache)))
        _cache[key] = Shard(file1, file2, num_states, backend)
    return _cache[key]


def clear():
    _cache.clear()
