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


def _alloc_rows(n, pitch, pinned):
    if pinned:
        import torch
        return torch.empty((n, pitch), dtype=torch.int8, pin_memory=True).numpy()     # the view keeps the pinned storage alive
    return np.empty((n, pitch), dtype=np.int8)


def _locations(names_raw, n_names, cid, starts, ends, n):
    uniq = names_raw.split(b"\0")[:n_names]
    table = np.array([u.decode() for u in uniq], dtype=object) if uniq else np.array([], dtype=object)
    return dict(chrom=table[cid] if n else np.array([], dtype=object), start=starts, end=ends,
                chrom_id=cid, chrom_names=b"\0".join(uniq) + b"\0")


def read_matrix(path, rows=None, want_locations=True, num_states=127, pinned=False, shape=None, split=None,
                return_total=False):
    """Parse one input matrix file: `chr start end state_1 ... state_C` (README.md:286-292).

    Returns (locations, states0): states0 is an int8 [rows, C] view (row pitch = multiple of 16 bytes, i.e. already
    in the kernels' layout; `states0.base` is the pitched buffer) holding label-1 (helpers.py:154-155); locations
    is None or dict(chrom=object array, start=int64 array, end=int64 array).
    `rows` = (lo, hi) restricts the parse to that row range (skiprows / nrows of helpers.py:154-155); `split` = (rank,
    world) selects this rank's range splitRows(total, world)[rank] instead, without knowing `total` beforehand.
    Labels outside 1..num_states raise (the reference would fail later with an IndexError).
    With `rows` and `shape` = (total rows, biosample columns) the file is read by the two-call packer (epi_pack_tsv);
    otherwise it is parsed in ONE pass (epi_tsv_parse_*: the row count comes out of the parse), which saves one of the
    two inflate passes that bound reading a gzipped matrix.  `return_total` appends the file's total row count."""
    path = Path(path)
    if not path.is_file():
        raise FileNotFoundError(str(path))
    from . import timing
    with timing.stage("inflate+parse"):
        return _read_matrix(path, rows, want_locations, num_states, pinned, shape, split, return_total)


def _read_matrix(path, rows, want_locations, num_states, pinned, shape, split, return_total):
    if rows is not None and shape is not None:
        total, cols = shape
        lo, hi = int(rows[0]), int(rows[1])
        n = max(hi - lo, 0)
        pitch = pitch_for(cols)
        buf = _alloc_rows(n, pitch, pinned)
        starts = np.empty(n, dtype=np.int64)
        ends = np.empty(n, dtype=np.int64)
        cid = np.empty(n, dtype=np.int32)
        names = ctypes.create_string_buffer(1 << 16)
        nnames = ctypes.c_int32(0)
        _lib.call("epi_pack_tsv", str(path).encode(), lo, hi, cols, int(num_states), ctypes.c_void_p(buf.ctypes.data),
                  pitch, ctypes.c_void_p(starts.ctypes.data), ctypes.c_void_p(ends.ctypes.data),
                  ctypes.c_void_p(cid.ctypes.data), names, len(names), ctypes.byref(nnames))
        names_raw, n_names = names.raw, nnames.value
    else:
        handle = ctypes.c_void_p(0)
        total_c, cols_c = ctypes.c_int64(0), ctypes.c_int32(0)
        nnames, nbytes = ctypes.c_int32(0), ctypes.c_int32(0)
        _lib.call("epi_tsv_parse_open", str(path).encode(), int(num_states), ctypes.byref(handle), ctypes.byref(total_c),
                  ctypes.byref(cols_c), ctypes.byref(nnames), ctypes.byref(nbytes))
        try:
            total, cols = total_c.value, cols_c.value
            if total > 0 and cols < 1:
                raise ValueError("%s: expected `chr start end state_1 ...` rows" % path)
            if split is not None:
                rows = splitRows(total, int(split[1]))[int(split[0])]
            lo, hi = (0, total) if rows is None else (int(rows[0]), int(rows[1]))
            n = max(hi - lo, 0)
            pitch = pitch_for(max(cols, 1))
            buf = _alloc_rows(n, pitch, pinned)
            starts = np.empty(n, dtype=np.int64)
            ends = np.empty(n, dtype=np.int64)
            cid = np.empty(n, dtype=np.int32)
            names = ctypes.create_string_buffer(max(nbytes.value, 1))
            _lib.call("epi_tsv_parse_fetch", handle, lo, hi, ctypes.c_void_p(buf.ctypes.data), pitch,
                      ctypes.c_void_p(starts.ctypes.data), ctypes.c_void_p(ends.ctypes.data),
                      ctypes.c_void_p(cid.ctypes.data), names, len(names))
            names_raw, n_names = names.raw, nnames.value
        finally:
            _lib.call("epi_tsv_parse_close", handle)
    states0 = buf[:, :max(cols, 0)]
    loc = _locations(names_raw, n_names, cid, starts, ends, n) if want_locations else None
    if return_total:
        return loc, states0, total
    return loc, states0


def readStates(file1Path=Path("null"), file2Path=Path("null"), rowsToCalc=(0, 0), expBool=True, verbose=True, groupSize=-1,
               numStates=127):
    """The reference's reader under its own name and return convention (helpers.py:123-194), for callers written
    against it: 0-based `int` arrays of rows [rowsToCalc[0], rowsToCalc[1]).
        single:                  file1Arr
        paired, expBool=True:    [file1Arr | file2Arr]
        paired, expBool=False:   file1Arr, file2Arr and the two halves of the per-row shuffle of [file1Arr | file2Arr]
                                 (np.argsort(np.random.rand(rows, N), axis=1): the same draws from numpy's global generator
                                 as the reference, so a seeded caller gets the reference's shuffle), cut at file1Arr's width
                                 or into two groups of groupSize.
    The stage drivers do not go through this shim (they keep the int8 matrix of read_matrix and shuffle on the device);
    `numStates` only bounds the label check of the native parser."""
    rows = (int(rowsToCalc[0]), int(rowsToCalc[1]))
    _, a = read_matrix(file1Path, rows, want_locations=False, num_states=numStates)
    file1Arr = a.astype(int)
    if str(file2Path) == "null":
        return file1Arr
    _, b = read_matrix(file2Path, rows, want_locations=False, num_states=numStates)
    file2Arr = b.astype(int)
    combinedArr = np.concatenate((file1Arr, file2Arr), axis=1)
    if expBool:
        return combinedArr
    randomIndices = np.argsort(np.random.rand(*combinedArr.shape), axis=1)
    shuffled = np.take_along_axis(combinedArr, randomIndices, axis=1)
    if groupSize == -1:
        return file1Arr, file2Arr, shuffled[:, :file1Arr.shape[1]], shuffled[:, file1Arr.shape[1]:]
    return file1Arr, file2Arr, shuffled[:, :groupSize], shuffled[:, groupSize:2 * groupSize]


def read_scores(path):
    """Parse a score file `chr start end score_1 .. score_K` (scores_*.txt.gz) with the native reader:
    returns (locations dict as read_matrix, float64 [rows, K] scores).  Replaces the pandas read of
    similaritySearch_max_mean.readScores (similaritySearch_max_mean.py:51-74); decimal text converts to the nearest
    double, which is what pandas' parser yields for these files."""
    path = Path(path)
    if not path.is_file():
        raise FileNotFoundError(str(path))
    handle = ctypes.c_void_p(0)
    rows_c, cols_c = ctypes.c_int64(0), ctypes.c_int32(0)
    nnames, nbytes = ctypes.c_int32(0), ctypes.c_int32(0)
    _lib.call("epi_scores_tsv_open", str(path).encode(), ctypes.byref(handle), ctypes.byref(rows_c), ctypes.byref(cols_c),
              ctypes.byref(nnames), ctypes.byref(nbytes))
    try:
        n, k = rows_c.value, max(cols_c.value, 0)
        scores = np.empty((n, k), dtype=np.float64)
        starts = np.empty(n, dtype=np.int64)
        ends = np.empty(n, dtype=np.int64)
        cid = np.empty(n, dtype=np.int32)
        names = ctypes.create_string_buffer(max(nbytes.value, 1))
        _lib.call("epi_scores_tsv_fetch", handle, ctypes.c_void_p(scores.ctypes.data), ctypes.c_void_p(starts.ctypes.data),
                  ctypes.c_void_p(ends.ctypes.data), ctypes.c_void_p(cid.ctypes.data), names, len(names))
    finally:
        _lib.call("epi_scores_tsv_close", handle)
    return _locations(names.raw, nnames.value, cid, starts, ends, n), scores


def savez_level(path, level=1, **arrays):
    """np.savez_compressed with a chosen deflate level: the same .npz container (np.load reads it), but level 1 instead of
    zlib's default 6 -- these score matrices are dominated by repeated quiescent rows and pack almost as well at a
    quarter of the time (whole chr1: temp_scores 0.38 s instead of 1.56 s, genome_stats 0.9 s instead of 3.2 s)."""
    import zipfile
    from . import timing
    path = Path(path)
    if path.suffix != ".npz":
        path = path.with_name(path.name + ".npz")
    with timing.stage("npz hand-over files"):
        with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED, allowZip64=True, compresslevel=level) as zf:
            for name, arr in arrays.items():
                with zf.open(name + ".npy", "w", force_zip64=True) as f:
                    np.lib.format.write_array(f, np.asanyarray(arr), allow_pickle=True)


def sharedToNumpy(sharedArr, numRows, numStates):
    """Kept for signature compatibility (helpers.py:315-327): view a flat float32 buffer as [rows, K]."""
    return np.frombuffer(sharedArr, dtype=np.float32).reshape((numRows, numStates))
