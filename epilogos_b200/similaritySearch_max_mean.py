"""Similarity-search preparation stage.  Mirror of similaritySearch_max_mean.py of the reference:

    main(outputDir, scoresPath, windowBins, blockSize, windowBP, filterState, filterScore)   (similaritySearch_max_mean.py:9)

reads a score file, selects non-overlapping salient windows with the max-mean rule (as many as fit the genome), cuts a
block-reduced slice of the scores around each, filters them, and block-reduces the whole genome.  It writes what the
distance engine (similaritySearch_calc) and the writer read, under the reference's names:

    genome_stats.npz      scores float64 [bins, K], coords object [bins, 3]                (:13)
    simsearch_cube.npz    scores float64 [regions, windowBins/blockSize, K], coords object [regions, 3]   (:42)
    reduced_genome.npy    float64 [ceil(bins / blockSize), K]                              (:158-160)

The text is parsed by the native reader (csrc/hostio.cu epi_scores_tsv_*), the window selection is the native
epi_roi_maxmean (csrc/roi.cu); the block reductions are a handful of vector operations over the score matrix.
"Reduction" keeps, of every block of blockSize consecutive bins, the bin whose score sum is largest.  The sums add the
states in order 0..K-1, which is what pandas computes for these frames.
"""
import sys
from pathlib import Path
from time import time

import numpy as np

from . import helpers, roi

SLICE_BATCH = 4096


def rowSums(stateScores):
    """Per-bin sum over the states, added left to right (pandas DataFrame.sum(axis=1) on a column-major block,
    similaritySearch_max_mean.py:67, 97, 152)."""
    total = np.zeros(stateScores.shape[0], dtype=np.float64)
    for s in range(stateScores.shape[1]):
        total += stateScores[:, s]
    return total


def readScores(scoresPath):
    """similaritySearch_max_mean.py:51-74.  Returns (locations dict, float64 [bins, K] scores, per-bin score sums)."""
    loc, scores = helpers.read_scores(scoresPath)
    return loc, scores, rowSums(scores)


def _coords(chrom, start, end):
    out = np.empty((len(chrom), 3), dtype=object)
    out[:, 0] = chrom
    out[:, 1] = np.asarray(start, dtype=np.int64).tolist()       # Python ints, as pandas' to_numpy() of a mixed frame yields
    out[:, 2] = np.asarray(end, dtype=np.int64).tolist()
    return out


def makeSlices(stateScores, sums, centers, windowBins, blockSize):
    """The cube of reduced windows (makeSlice, similaritySearch_max_mean.py:77-98): the window around `center` is cut
    into consecutive blocks of blockSize bins and the FIRST bin with the largest score sum of each block is kept
    (idxmax)."""
    centers = np.asarray(centers, dtype=np.int64)
    half = windowBins // 2
    width = 2 * half + (1 if windowBins % 2 else 0)
    n_blocks = -(-width // blockSize)
    pad = n_blocks * blockSize - width
    k = stateScores.shape[1]
    cube = np.empty((len(centers), n_blocks, k), dtype=np.float64)
    offs = np.arange(width, dtype=np.int64)
    for b0 in range(0, len(centers), SLICE_BATCH):
        c = centers[b0:b0 + SLICE_BATCH]
        rows = (c - half)[:, None] + offs[None, :]                      # [r, width] genome rows of each window
        if rows.size and (rows.min() < 0 or rows.max() >= len(sums)):
            raise IndexError("similarity-search window leaves the genome")
        w = sums[rows]
        if pad:
            w = np.concatenate((w, np.full((len(c), pad), -np.inf)), axis=1)
        best = np.argmax(w.reshape(len(c), n_blocks, blockSize), axis=2)        # first maximum
        pick = rows[np.arange(len(c))[:, None], np.arange(n_blocks)[None, :] * blockSize + best]
        cube[b0:b0 + len(c)] = stateScores[pick]
    return cube


def removeRegions(roiCoords, roiCube, filterState, filterScore):
    """similaritySearch_max_mean.py:101-134: drop windows over two chromosomes, windows whose strongest state is
    filterState (1-based; -1 = the last state; 0 = no filter), and windows whose largest score is below filterScore
    (-1 = no filter)."""
    drop = np.zeros(len(roiCube), dtype=bool)
    if len(roiCube):
        drop |= np.asarray(roiCoords[:, 1], dtype=np.int64) >= np.asarray(roiCoords[:, 2], dtype=np.int64)
        if filterState != 0:
            fs = roiCube.shape[2] - 1 if filterState == -1 else filterState - 1
            drop |= np.argmax(np.max(roiCube, axis=1), axis=1) == fs
        if filterScore != -1:
            drop |= np.max(roiCube, axis=(1, 2)) < filterScore
    return roiCoords[~drop], roiCube[~drop]


def reducedIndices(sums, blockSize):
    """Row kept of every block of the genome (reduceGenome, similaritySearch_max_mean.py:137-160): the bin with the
    largest score sum; among equal sums the reference keeps whichever its unstable sort puts last -- here the last bin
    of the block with that sum (bins with equal sums are almost always equal rows, and then the choice is immaterial)."""
    n = len(sums)
    n_blocks = -(-n // blockSize)
    padded = np.full(n_blocks * blockSize, -np.inf)
    padded[:n] = sums
    blocks = padded.reshape(n_blocks, blockSize)
    last = blockSize - 1 - np.argmax(blocks[:, ::-1], axis=1)
    return np.arange(n_blocks, dtype=np.int64) * blockSize + last


def reduceGenome(outputDir, stateScores, sums, blockSize):
    reduced = stateScores[reducedIndices(sums, blockSize)]
    np.save(Path(outputDir) / "reduced_genome.npy", reduced, allow_pickle=True)
    return reduced


def main(outputDir, scoresPath, windowBins, blockSize, windowBP, filterState, filterScore):
    outputDir = Path(outputDir)
    print("Reading in data...", flush=True); t = time()
    loc, stateScores, sums = readScores(scoresPath)
    helpers.savez_level(outputDir / "genome_stats", scores=stateScores, coords=_coords(loc["chrom"], loc["start"], loc["end"]))
    # as many regions as could tile the genome (similaritySearch_max_mean.py:15-17)
    maxRegions = int(stateScores.shape[0] // windowBins)
    print("    Time:", format(time() - t, '.0f'), "seconds\n", flush=True)

    print("Finding regions of size {}kb...".format(windowBP // 1000), flush=True); t1 = time()
    sel = roi.max_mean(loc["start"], loc["end"], sums, windowBins, maxRegions)
    centers = sel["original_idx"]
    roiCoords = _coords(loc["chrom"][centers], sel["start"], sel["end"])
    print("    Time:", format(time() - t1, '.0f'), "seconds\n", flush=True)

    print("Reducing region scores by factor of {}...".format(blockSize), flush=True); t1 = time()
    roiCube = makeSlices(stateScores, sums, centers, windowBins, blockSize)
    print("    Time:", format(time() - t1, '.0f'), "seconds\n", flush=True)

    print("Filtering out uninteresting regions...", flush=True); t1 = time()
    roiCoords, roiCube = removeRegions(roiCoords, roiCube, filterState, filterScore)
    helpers.savez_level(outputDir / "simsearch_cube", scores=roiCube, coords=roiCoords)
    print("    Time:", format(time() - t1, '.0f'), "seconds\n", flush=True)

    print("Reducing genome scores by factor of {}...".format(blockSize), flush=True); t1 = time()
    reduceGenome(outputDir, stateScores, sums, blockSize)
    print("    Time:", format(time() - t1, '.0f'), "seconds\n", flush=True)
    print("Total time:", format(time() - t, '.0f'), "seconds\n", flush=True)


if __name__ == "__main__":
    main(Path(sys.argv[1]), Path(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]),
         float(sys.argv[7]))
