"""Text / npz writers of the score stage (scores.py:509-536, 163-169)."""
import gzip

import numpy as np


def write_scores_text(path, scores32, loc):
    """`chr \\t start \\t end \\t K x "{:.5f}"` per bin through gzip (scores.py:530-536).  The reference formats
    np.float32 values with "{:.5f}", i.e. the exact binary value rounded to 5 decimals -- identical to
    "%.5f" % float(v)."""
    scores32 = np.asarray(scores32, dtype=np.float32)
    rows, k = scores32.shape
    fmt = "\t".join(["%.5f"] * k)
    chrom, start, end = loc["chrom"], loc["start"], loc["end"]
    with gzip.open(path, "wt") as out:
        step = 65536
        for lo in range(0, rows, step):
            hi = min(rows, lo + step)
            block = scores32[lo:hi].astype(np.float64)
            lines = ["%s\t%d\t%d\t%s\n" % (chrom[i], start[i], end[i], fmt % tuple(block[i - lo]))
                     for i in range(lo, hi)]
            out.write("".join(lines))


def location_array(loc):
    """object [rows, 3] array as produced by pd.read_table(...).to_numpy() in scores.py:161."""
    arr = np.empty((len(loc["chrom"]), 3), dtype=object)
    arr[:, 0] = loc["chrom"]
    arr[:, 1] = loc["start"]
    arr[:, 2] = loc["end"]
    return arr
