"""Similarity-search output stage.  Mirror of similaritySearch_write.py of the reference:

    main(outputDir, windowBins, blockSize, nJobs, nDesiredMatches)                       (similaritySearch_write.py:14)

gathers the `simsearch_indices_<job>.npy` files of the distance engine, turns reduced-genome indices into genomic
coordinates and writes `simsearch.bed.gz`: one line per region, `chrom \\t start \\t end \\t JSON list of "chrom:start:end"`
(the region itself first, then its matches), sorted by (chrom, start); then removes `genome_stats.npz` and the per-job
files and stores the combined `simsearch_indices.npy` (:175-188).

The reference compresses with pysam's bgzip and adds a tabix index (:161-167).  pysam / htslib are not part of this
image, so `simsearch.bed.gz` is written here as BGZF (the blocked gzip dialect bgzip produces: any gzip reader and tabix
itself accept it) and the `.tbi` index is NOT produced -- `tabix -p bed simsearch.bed.gz` creates it where htslib is
installed.  The query mode of `simsearch` (similaritySearch_run.querySimSearch) reads the bed file only.
"""
import json
import os
import struct
import sys
import zlib
from pathlib import Path
from time import time

import numpy as np

from .helpers import splitRows

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_BGZF_BLOCK = 0xff00


def bgzf_write(path, data, level=6):
    """`data` (bytes) as a BGZF file: gzip members of at most 64 KiB of input, each with the `BC` extra field holding
    the member's size, closed by the empty end-of-file member (SAM specification, section 4.1)."""
    with open(path, "wb") as f:
        for off in range(0, len(data), _BGZF_BLOCK):
            chunk = data[off:off + _BGZF_BLOCK]
            comp = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = comp.compress(chunk) + comp.flush()
            bsize = 12 + 6 + len(body) + 8 - 1                         # header + extra field + deflate + crc/isize, minus one
            f.write(struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 0x42, 0x43, 2, bsize))
            f.write(body)
            f.write(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        f.write(_BGZF_EOF)


def reduceGenomeCoords(inputDir, blockSize):
    """Coordinates of the reduced genome (similaritySearch_write.py:44-67): chromosome and start of the first bin and
    end of the last bin of every block of blockSize bins.  Returns an object array [blocks, 3]."""
    coords = np.load(Path(inputDir) / "genome_stats.npz", allow_pickle=True)["coords"]
    n = len(coords)
    first = np.arange(0, n, blockSize)
    last = np.minimum(first + blockSize - 1, n - 1)
    out = np.empty((len(first), 3), dtype=object)
    out[:, 0] = coords[first, 0]
    out[:, 1] = coords[first, 1]
    out[:, 2] = coords[last, 2]
    return out


def readSimsearchIndices(inputDir, nRegions, nDesiredMatches, nJobs):
    """similaritySearch_write.py:70-92."""
    arr = np.zeros((nRegions, nDesiredMatches), dtype=np.int32)
    rowList = splitRows(nRegions, nJobs)
    for file in Path(inputDir).glob("simsearch_indices_*.npy"):
        i = int(file.stem.split("_")[-1])
        arr[rowList[i][0]:rowList[i][1]] = np.load(file, allow_pickle=True)
    return arr


def convertIndicesToCoords(simsearchArr, reducedGenomeCoords, roiCoords, windowBins, blockSize, nRegions, nDesiredMatches):
    """similaritySearch_write.py:95-124: object array [regions, 1 + nDesiredMatches, 3], the query region first.
    Entries of -1 index from the end, as in the reference; the writer skips them."""
    flat = simsearchArr.reshape(-1).astype(np.int64)
    res = np.empty((nRegions * nDesiredMatches, 3), dtype=object)
    res[:, :2] = reducedGenomeCoords[flat, :2]
    res[:, 2] = reducedGenomeCoords[flat + windowBins // blockSize - 1, 2]
    res = res.reshape(nRegions, nDesiredMatches, 3)
    return np.concatenate((np.asarray(roiCoords, dtype=object).reshape(nRegions, 1, 3), res), axis=1)


def resultLines(searchResults, simsearchArr, roiCoords):
    """The text of simsearch.bed.gz before compression (similaritySearch_write.py:142-160)."""
    rows = []
    for r in range(len(simsearchArr)):
        keep = np.concatenate(([True], simsearchArr[r] != -1))
        recs = ["{}:{}:{}".format(c, s, e) for c, s, e in searchResults[r][keep]]
        rows.append((roiCoords[r][0], roiCoords[r][1], roiCoords[r][2], json.dumps(recs)))
    order = sorted(range(len(rows)), key=lambda i: (rows[i][0], rows[i][1]))          # stable, like sort_values
    return "".join("{}\t{}\t{}\t{}\n".format(*rows[i]) for i in order)


def writeResults(outputDir, searchResults, simsearchArr, roiCoords, nRegions):
    fn = os.path.join(outputDir, "simsearch.bed.gz")
    for stale in (fn, fn + ".tbi"):
        if os.path.exists(stale):
            os.remove(stale)
    bgzf_write(fn, resultLines(searchResults, simsearchArr, roiCoords).encode())
    if not os.path.exists(fn) or os.stat(fn).st_size == 0:
        raise Exception("Error: Could not create bgzip archive [{}]".format(fn))


def cleanUpFiles(outputDir, simsearchArr):
    """similaritySearch_write.py:175-188."""
    outputDir = Path(outputDir)
    os.remove(outputDir / "genome_stats.npz")
    for file in outputDir.glob("simsearch_indices_*.npy"):
        os.remove(file)
    np.save(outputDir / "simsearch_indices.npy", simsearchArr, allow_pickle=True)


def main(outputDir, windowBins, blockSize, nJobs, nDesiredMatches):
    outputDir = Path(outputDir)
    print("Reducing genome coordinates...", flush=True); t = time()
    reducedGenomeCoords = reduceGenomeCoords(outputDir, blockSize)
    print("    Time:", format(time() - t, '.0f'), "seconds\n", flush=True)

    print("Reading in search results...", flush=True); t1 = time()
    cube = np.load(outputDir / "simsearch_cube.npz", allow_pickle=True)
    nRegions = cube["scores"].shape[0]
    roiCoords = cube["coords"]
    simsearchArr = readSimsearchIndices(outputDir, nRegions, nDesiredMatches, nJobs)
    searchResults = convertIndicesToCoords(simsearchArr, reducedGenomeCoords, roiCoords, windowBins, blockSize, nRegions,
                                           nDesiredMatches)
    print("    Time:", format(time() - t1, '.0f'), "seconds\n", flush=True)

    print("Writing search results...", flush=True); t1 = time()
    writeResults(outputDir, searchResults, simsearchArr, roiCoords, nRegions)
    print("    Time:", format(time() - t1, '.0f'), "seconds\n", flush=True)

    print("Cleaning up temp files...", flush=True)
    cleanUpFiles(outputDir, simsearchArr)
    print("Total time:", format(time() - t, '.0f'), "seconds\n", flush=True)


if __name__ == "__main__":
    main(Path(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
