"""Host-side helpers of the scoring path (mirror of the hot subset of the reference's helpers.py).

readStates of the reference (helpers.py:123-194) returns an int64 ndarray of labels-1; here the same rows
are packed straight into the int8 [bins, pitch] layout the kernels consume (see engine.pack_states).
"""
import gzip
from pathlib import Path

import numpy as np


def strToBool(string):
    """'True' / 'False' -> bool (helpers.py:47-60); anything else raises ValueError."""
    if string == "True":
        return True
    if string == "False":
        return False
    raise ValueError("Invalid boolean string")


def _open(path):
    path = Path(path)
    return gzip.open(path, "rb") if path.name.endswith("gz") else open(path, "rb")


def countRows(dataFilePath):
    """Number of newline characters in the (possibly gzipped) file (helpers.py:80-99)."""
    total = 0
    with _open(dataFilePath) as f:
        while True:
            block = f.read(1 << 20)
            if not block:
                break
            total += block.count(b"\n")
    return total


def splitRows(totalRows, numProcesses):
    """Row ranges [i*T//n, (i+1)*T//n) -- one per worker in the reference (helpers.py:102-120), one per
    GPU rank here."""
    return [(i * totalRows // numProcesses, (i + 1) * totalRows // numProcesses) for i in range(numProcesses)]


def read_matrix(path, rows=None, want_locations=True):
    """Parse one input matrix file: `chr start end state_1 ... state_C` (README.md:286-292).

    Returns (locations, states0) where states0 is int8 [rows, C] holding label-1 (helpers.py:154-155) and
    locations is None or a dict(chrom=object array, start=int64 array, end=int64 array).
    `rows` = (lo, hi) restricts the parse to that row range (skiprows / nrows of helpers.py:154-155).
    """
    import pandas as pd
    path = Path(path)
    kw = dict(header=None, sep="\t")
    if rows is not None:
        kw.update(skiprows=rows[0], nrows=rows[1] - rows[0])
    ncols = pd.read_table(path, nrows=1, header=None, sep="\t").shape[1]
    states = pd.read_table(path, usecols=range(3, ncols), dtype=np.int16, **kw).to_numpy()
    if states.size and (states.min() < 1 or states.max() > 127):
        raise ValueError("%s: state labels must be integers in [1, 127]" % path)
    states0 = (states - 1).astype(np.int8)
    loc = None
    if want_locations:
        df = pd.read_table(path, usecols=[0, 1, 2], **kw)
        loc = dict(chrom=df[0].to_numpy(dtype=object), start=df[1].to_numpy(dtype=np.int64),
                   end=df[2].to_numpy(dtype=np.int64))
    return loc, states0


def sharedToNumpy(sharedArr, numRows, numStates):
    """Kept for signature compatibility (helpers.py:315-327): view a flat float32 buffer as [rows, K]."""
    return np.frombuffer(sharedArr, dtype=np.float32).reshape((numRows, numStates))
