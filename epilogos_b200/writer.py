"""Text / npz writers of the score stage (scores.py:509-536, 163-169)."""
import ctypes

import numpy as np

from . import _lib


def write_scores_text(path, scores32, loc, level=None, threads=0):
    """`chr \\t start \\t end \\t K x "{:.5f}"` per bin through gzip (scores.py:530-536), formatted and compressed by
    the native writer (epi_write_scores_gz).  The reference formats np.float32 values with "{:.5f}", i.e. the exact
    binary value rounded to 5 decimals, which is what printf's "%.5f" of the widened double gives.
    `level`: gzip level, default $EPILOGOS_B200_GZIP_LEVEL or 4.  Measured on 1 M rows x 18 scores, 8 vCPUs: level 9 (what
    the reference writes) 0.15 M rows/s, 40.2 MB; 6: 0.86 M rows/s, 41.6 MB; 4: 1.85 M rows/s, 44.4 MB; 1: 2.45 M rows/s,
    51.1 MB -- the decompressed text is the same at any level."""
    if level is None:
        import os
        level = int(os.environ.get("EPILOGOS_B200_GZIP_LEVEL", "4"))
    scores32 = np.ascontiguousarray(scores32, dtype=np.float32)
    rows, k = scores32.shape
    if "chrom_id" in loc:
        cid = np.ascontiguousarray(loc["chrom_id"], dtype=np.int32)
        names = loc["chrom_names"]
    else:
        uniq, cid = np.unique(np.asarray(loc["chrom"], dtype=object).astype(str), return_inverse=True)
        cid = cid.astype(np.int32)
        names = b"\0".join(u.encode() for u in uniq) + b"\0"
    if rows == 0:
        names = names or b"\0"
    starts = np.ascontiguousarray(loc["start"], dtype=np.int64)
    ends = np.ascontiguousarray(loc["end"], dtype=np.int64)
    from . import timing
    with timing.stage("format %.5f + deflate"):
        _lib.call("epi_write_scores_gz", str(path).encode(), ctypes.c_char_p(names), ctypes.c_void_p(cid.ctypes.data),
                  ctypes.c_void_p(starts.ctypes.data), ctypes.c_void_p(ends.ctypes.data),
                  ctypes.c_void_p(scores32.ctypes.data), rows, k, int(level), int(threads))


def location_array(loc):
    """object [rows, 3] array as produced by pd.read_table(...).to_numpy() in scores.py:161."""
    arr = np.empty((len(loc["chrom"]), 3), dtype=object)
    arr[:, 0] = loc["chrom"]
    arr[:, 1] = loc["start"]
    arr[:, 2] = loc["end"]
    return arr
