"""Host-side helpers of the scoring path (mirror of the hot subset of the reference's helpers.py).

readStates of the reference (helpers.py:123-194) returns an int64 ndarray of labels-1 parsed with pandas; here
the rows are parsed by the library's native packer (epi_pack_tsv) straight into the pinned int8 [bins, pitch]
layout the kernels consume.
"""
import ctypes
from pathlib import Path

import numpy as np

from . import _lib


def strToBool(string):
    """'True' / 'False' -> bool (helpers.py:47-60); anything else raises ValueError."""
    if string == "True":
        return True
    if string == "False":
        return False
    raise ValueError("Invalid boolean string")


def tsv_shape(path):
    """(rows, biosample columns): rows = newline count as in helpers.countRows (helpers.py:80-99)."""
    rows, cols = ctypes.c_int64(0), ctypes.c_int32(0)
    _lib.call("epi_tsv_shape", str(path).encode(), ctypes.byref(rows), ctypes.byref(cols))
    return rows.value, cols.value


def countRows(dataFilePath):
    """Number of newline characters in the (possibly gzipped) file (helpers.py:80-99)."""
    p = Path(dataFilePath)
    if not p.is_file():
        raise FileNotFoundError(str(p))
    return tsv_shape(p)[0]


def splitRows(totalRows, numProcesses):
    """Row ranges [i*T//n, (i+1)*T//n) -- one per worker in the reference (helpers.py:102-120), one per
    GPU rank here."""
    return [(i * totalRows // numProcesses, (i + 1) * totalRows // numProcesses) for i in range(numProcesses)]


def pitch_for(cols):
    return (int(cols) + 15) & ~15


def read_matrix(path, rows=None, want_locations=True, num_states=127, pinned=False, shape=None):
    """Parse one input matrix file: `chr start end state_1 ... state_C` (README.md:286-292).

    Returns (locations, states0): states0 is an int8 [rows, C] view (row pitch = multiple of 16 bytes, i.e. already
    in the kernels' layout; `states0.base` is the pitched buffer) holding label-1 (helpers.py:154-155); locations
    is None or dict(chrom=object array, start=int64 array, end=int64 array).
    `rows` = (lo, hi) restricts the parse to that row range (skiprows / nrows of helpers.py:154-155).
    Labels outside 1..num_states raise (the reference would fail later with an IndexError).
    `shape` = (total rows, biosample columns) if the caller already knows them (saves one inflate pass over the file)."""
    path = Path(path)
    if not path.is_file():
        raise FileNotFoundError(str(path))
    total, cols = tsv_shape(path) if shape is None else shape
    if cols < 1:
        raise ValueError("%s: expected `chr start end state_1 ...` rows" % path)
    lo, hi = (0, total) if rows is None else (int(rows[0]), int(rows[1]))
    n = max(hi - lo, 0)
    pitch = pitch_for(cols)
    if pinned:
        import torch
        holder = torch.empty((n, pitch), dtype=torch.int8, pin_memory=True)
        buf = holder.numpy()
    else:
        holder = None
        buf = np.empty((n, pitch), dtype=np.int8)
    starts = np.empty(n, dtype=np.int64)
    ends = np.empty(n, dtype=np.int64)
    cid = np.empty(n, dtype=np.int32)
    names = ctypes.create_string_buffer(1 << 16)
    nnames = ctypes.c_int32(0)
    _lib.call("epi_pack_tsv", str(path).encode(), lo, hi, cols, int(num_states), ctypes.c_void_p(buf.ctypes.data),
              pitch, ctypes.c_void_p(starts.ctypes.data), ctypes.c_void_p(ends.ctypes.data),
              ctypes.c_void_p(cid.ctypes.data), names, len(names), ctypes.byref(nnames))
    states0 = buf[:, :cols]
    loc = None
    if want_locations:
        uniq = names.raw.split(b"\0")[: nnames.value]
        table = np.array([u.decode() for u in uniq], dtype=object) if uniq else np.array([], dtype=object)
        loc = dict(chrom=table[cid] if n else np.array([], dtype=object), start=starts, end=ends,
                   chrom_id=cid, chrom_names=b"\0".join(uniq) + b"\0")
    return loc, states0          # states0 (a numpy view) keeps the pinned torch storage alive


def sharedToNumpy(sharedArr, numRows, numStates):
    """Kept for signature compatibility (helpers.py:315-327): view a flat float32 buffer as [rows, K]."""
    return np.frombuffer(sharedArr, dtype=np.float32).reshape((numRows, numStates))
