"""ChromHMM state calls -> Epilogos input matrices.  Mirror of the reference's bin/preprocess_data_ChromHMM.sh.

    python -m epilogos_b200.preprocess <datadir> <metadata> <chromsizes> [-o OUTDIR] [-z LEVEL] [-j STATES]

The script (preprocess_data_ChromHMM.sh:34-49) walks the chromosomes of <chromsizes>, collects for every biosample of
<metadata> (column 1, header skipped) the file `<datadir>/*<biosample>*<chr>_*.txt*` that ChromHMM's `-printstatebyline`
wrote (line 1 `<biosample> <chr>`, line 2 `MaxState E`, then one state label per 200 bp bin), pastes them side by side and
prefixes every row with `chr, start, end`: `matrix_<chr>.txt`, the input format of `epilogos -i` (README.md:286-292).

Here every file is inflated and parsed natively straight into its COLUMN of the int8 bins x biosamples matrix (several files
at a time, csrc/hostio.cu: epi_statebyline_read) -- `read_chromosome` hands that matrix to a caller that wants to score it
without ever writing the text -- and `main` writes the same `matrix_<chr>.txt` files as the script, byte for byte
(epi_write_matrix_tsv; `-z LEVEL` gzips them, which `epilogos -i` reads as well).  Same messages on stdout.
Differences: biosample files of unequal length are an error (paste would pad the short ones with empty fields, which no later
stage accepts); more than one file matching a biosample and chromosome is an error (the script's command line breaks).
"""
import ctypes
import glob
import os
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from . import _lib, helpers

BIN_SIZE = 200          # preprocess_data_ChromHMM.sh:47


def biosamples(metadata):
    """Column 1 of the metadata file without its header line (`cut -f1 | tail -n +2`)."""
    with open(metadata) as f:
        lines = f.read().splitlines()
    return [line.split("\t")[0] for line in lines[1:] if line != ""]


def chromosomes(chromsizes):
    with open(chromsizes) as f:
        return [line.split("\t")[0] for line in f.read().splitlines() if line != ""]


def find_files(datadir, names, chrom):
    """For every biosample the file `<datadir>/*<biosample>*<chr>_*.txt*`, in metadata order; biosamples without one are
    skipped (preprocess_data_ChromHMM.sh:38-43)."""
    found = []
    for name in names:
        hits = sorted(glob.glob(os.path.join(str(datadir), "*%s*%s_*.txt*" % (glob.escape(name), glob.escape(chrom)))))
        if len(hits) > 1:
            raise ValueError("more than one file for biosample %s on %s: %s" % (name, chrom, ", ".join(hits)))
        if hits:
            found.append(hits[0])
    return found


def _read_column(path, column, stride, cap, num_states):
    rows = ctypes.c_int64(0)
    chrom = ctypes.create_string_buffer(256)
    _lib.call("epi_statebyline_read", str(path).encode(), ctypes.c_void_p(column), stride, cap, int(num_states),
              ctypes.byref(rows), chrom, len(chrom))
    return rows.value, chrom.value.decode()


def read_chromosome(files, num_states=127, pinned=False, workers=None):
    """The int8 [bins, biosamples] matrix (label - 1; a view of a pitched buffer, as helpers.read_matrix returns it) of one
    chromosome from its per-biosample state-by-line files, and the chromosome name the files carry."""
    files = list(files)
    if not files:
        raise ValueError("no files")
    bins, chrom = _read_column(files[0], 0, 1, 0, num_states)               # first file: count the bins
    pitch = helpers.pitch_for(len(files))
    buf = helpers._alloc_rows(bins, pitch, pinned)
    workers = workers or max(1, min(len(files), (os.cpu_count() or 2)))
    # every file is parsed into a contiguous column (a byte per bin written 848 bytes apart straight into the matrix costs
    # three times the parse), groups of columns are then turned into rows by a blocked transpose
    group = 256
    cols_buf = np.empty((min(group, len(files)), max(bins, 1)), dtype=np.int8)
    names = []

    def job(j, slot):
        rows, c = _read_column(files[j], cols_buf.ctypes.data + slot * cols_buf.strides[0], 1, bins, num_states)
        if rows != bins:
            raise ValueError("%s holds %d bins, %s holds %d" % (files[j], rows, files[0], bins))
        return c
    _lib.call("epi_reader_concurrency", int(workers), 0)
    try:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            for g0 in range(0, len(files), group):
                g1 = min(len(files), g0 + group)
                names += list(pool.map(job, range(g0, g1), range(g1 - g0)))
                _lib.call("epi_columns_to_rows", ctypes.c_void_p(cols_buf.ctypes.data), int(cols_buf.strides[0]), g1 - g0, bins,
                          ctypes.c_void_p(buf.ctypes.data + g0), pitch, 0)
    finally:
        _lib.call("epi_reader_concurrency", 0, 0)
    buf[:, len(files):] = 0
    # awk takes the chromosome from the first pasted line, i.e. from the first file (preprocess_data_ChromHMM.sh:47)
    return buf[:, :len(files)], names[0] if names else chrom


def write_matrix(path, chrom, states0, gzip_level=None, threads=0, first_bin=0):
    """`chr start end label_1 .. label_C` rows (README.md:286-292) for an int8 label-1 matrix; gzip_level None = plain text."""
    states0 = np.asarray(states0)
    if states0.dtype != np.int8:
        states0 = states0.astype(np.int8)
    if states0.ndim != 2:
        raise ValueError("a [bins, biosamples] matrix is expected")
    if states0.strides[1] != 1:
        states0 = np.ascontiguousarray(states0)
    rows, cols = states0.shape
    _lib.call("epi_write_matrix_tsv", str(path).encode(), str(chrom).encode(), ctypes.c_void_p(states0.ctypes.data), rows, cols,
              int(states0.strides[0]) if rows > 1 else max(cols, 1), BIN_SIZE, int(first_bin),
              -1 if gzip_level is None else int(gzip_level), int(threads))


def main(datadir, metadata, chromsizes, outdir=".", gzip_level=None, num_states=127, out=sys.stdout):
    """The script's loop (preprocess_data_ChromHMM.sh:34-56) with its messages; returns the files written."""
    names = biosamples(metadata)
    written = []
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    for chrom in chromosomes(chromsizes):
        out.write("Processing %s: " % chrom)
        files = find_files(datadir, names, chrom)
        out.write("%d files found. " % len(files))
        if files:
            states0, named = read_chromosome(files, num_states)
            path = outdir / ("matrix_%s.txt%s" % (chrom, "" if gzip_level is None else ".gz"))
            write_matrix(path, named, states0, gzip_level)
            written.append(path)
            out.write("Done.\n")
        else:
            out.write("Skipping.\n")
        out.flush()
    return written


def _cli(argv):
    import argparse
    ap = argparse.ArgumentParser(prog="python -m epilogos_b200.preprocess", description=__doc__.split("\n\n")[0])
    ap.add_argument("datadir")
    ap.add_argument("metadata")
    ap.add_argument("chromsizes")
    ap.add_argument("-o", "--outdir", default=".")
    ap.add_argument("-z", "--gzip", type=int, default=None, metavar="LEVEL", help="write matrix_<chr>.txt.gz at this gzip level")
    ap.add_argument("-j", "--num-states", type=int, default=127, help="labels above this are an error (default: 127)")
    a = ap.parse_args(argv)
    main(a.datadir, a.metadata, a.chromsizes, a.outdir, a.gzip, a.num_states)


if __name__ == "__main__":
    _cli(sys.argv[1:])
