"""Similarity-search distance engine.  Mirror of similaritySearch_calc.py of the reference (main :13-34,
runEuclideanDistance :67-123, euclideanDistanceMulti :126-181): for every region of interest of `simsearch_cube.npz` the
squared Euclidean distance to every window of `reduced_genome.npy`, half the mode of those distances as the acceptance
threshold, and a greedy pick of up to nDesiredMatches non-overlapping windows in increasing distance.  Writes
`simsearch_indices_<processTag>.npy` (int32 [regions, nDesiredMatches]; -1 after a threshold stop) like the reference.

Everything runs on the GPU per batch of ROIs (csrc/simsearch.cu): distances, the sort (torch.sort = a library radix
sort, plumbing), the mode read off the sorted rows, and the greedy pick (one warp per ROI); only the int32 result rows
come back.  `nCores` is accepted and ignored.  Ties: the reference visits exactly tied distances in the order of numpy's
unstable introsort; here ties are visited in ascending window index.
"""
import ctypes
import sys
from pathlib import Path

import numpy as np

from . import _lib
from .helpers import splitRows

ROI_BATCH = 8


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def window_distances(genome_dev, xx_dev, rois_dev):
    """float64 CUDA tensors: genome [G, K], its row norms [G], ROIs [R, nS, K] -> distances [R, G - nS + 1]."""
    import torch
    g, k = genome_dev.shape
    r, ns, _ = rois_dev.shape
    out = torch.empty((r, g - ns + 1), dtype=torch.float64, device=genome_dev.device)
    _lib.call("epi_simsearch_distances", _ptr(genome_dev), _ptr(xx_dev), g, k, _ptr(rois_dev), r, ns, _ptr(out), _stream())
    return out


def row_norms(genome_dev):
    import torch
    xx = torch.empty(genome_dev.shape[0], dtype=torch.float64, device=genome_dev.device)
    _lib.call("epi_simsearch_row_norms", _ptr(genome_dev), genome_dev.shape[0], genome_dev.shape[1], _ptr(xx), _stream())
    return xx


def mode_of_sorted(sorted_dev):
    """scipy.stats.mode of every row of an ascending [R, W] float64 CUDA tensor (smallest value among ties)."""
    import torch
    r, w = sorted_dev.shape
    mode = torch.empty(r, dtype=torch.float64, device=sorted_dev.device)
    _lib.call("epi_simsearch_mode_sorted", _ptr(sorted_dev), r, w, _ptr(mode), ctypes.c_void_p(0), _stream())
    return mode


def pick(sorted_dev, index_dev, mode_dev, region_start_dev, n_super, n_desired):
    """The greedy selection (similaritySearch_calc.py:103-123) for every row: int32 CUDA tensor [R, n_desired]."""
    import torch
    r, w = sorted_dev.shape
    out = torch.empty((r, n_desired), dtype=torch.int32, device=sorted_dev.device)
    _lib.call("epi_simsearch_pick", _ptr(sorted_dev), _ptr(index_dev.contiguous()), r, w, _ptr(mode_dev),
              _ptr(region_start_dev.contiguous()), int(n_super), int(n_desired), _ptr(out), _stream())
    return out


def _region_starts(genome_coords, roi_coords, block_size):
    """Row of the non-reduced genome where each ROI starts, // blockSize (similaritySearch_calc.py:107-109: the FIRST row
    whose chromosome and start match).  The genome is cut into runs of equal chromosome once (O(rows)); a ROI is looked up
    by binary search inside the runs of its chromosome when their starts ascend, by a scan otherwise."""
    chrom = np.asarray(genome_coords[:, 0]).astype(str)
    start = np.asarray(genome_coords[:, 1]).astype(np.int64)
    n = len(chrom)
    cuts = np.flatnonzero(chrom[1:] != chrom[:-1]) + 1 if n > 1 else np.array([], dtype=np.int64)
    bounds = np.concatenate(([0], cuts, [n])).astype(np.int64)
    runs = {}
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi > lo:
            seg = start[lo:hi]
            runs.setdefault(chrom[lo], []).append((int(lo), int(hi), bool(np.all(seg[1:] >= seg[:-1]))))
    out = np.empty(len(roi_coords), dtype=np.int64)
    for r in range(len(roi_coords)):
        c, s = str(roi_coords[r, 0]), int(roi_coords[r, 1])
        found = -1
        for lo, hi, ascending in runs.get(c, ()):
            if ascending:
                p = lo + int(np.searchsorted(start[lo:hi], s, side="left"))
                if p < hi and start[p] == s:
                    found = p
            else:
                hits = np.flatnonzero(start[lo:hi] == s)
                if len(hits):
                    found = lo + int(hits[0])
            if found >= 0:
                break
        if found < 0:
            raise IndexError("region %s:%d is not a bin of genome_stats.npz" % (c, s))
        out[r] = found // block_size
    return out


def euclideanDistanceMulti(outputDir, genomeCoords, roiCoords, roiCube, windowBins, blockSize, nCores, nDesiredMatches,
                           rowsToCalc, processTag):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("epilogos_b200 has no CPU path: similarity search needs a CUDA device")
    outputDir = Path(outputDir)
    reduced = np.load(outputDir / "reduced_genome.npy", allow_pickle=True).astype(np.float64)
    lo, hi = rowsToCalc
    n_regions = hi - lo
    n_super = windowBins // blockSize
    result = np.zeros((n_regions, nDesiredMatches), dtype=np.int32)
    if n_regions == 0:
        np.save(outputDir / "simsearch_indices_{}.npy".format(processTag), result, allow_pickle=True)
        return result
    starts = _region_starts(np.asarray(genomeCoords), np.asarray(roiCoords)[lo:hi], blockSize)
    genome = torch.from_numpy(np.ascontiguousarray(reduced)).cuda()
    xx = row_norms(genome)
    cube = np.ascontiguousarray(np.asarray(roiCube)[lo:hi], dtype=np.float64)
    starts_dev = torch.from_numpy(starts).cuda()
    for b0 in range(0, n_regions, ROI_BATCH):
        rois = torch.from_numpy(cube[b0:b0 + ROI_BATCH]).cuda()
        dist = window_distances(genome, xx, rois)
        svals, sidx = torch.sort(dist, dim=1, stable=True)
        mode = mode_of_sorted(svals)
        picks = pick(svals, sidx, mode, starts_dev[b0:b0 + rois.shape[0]], n_super, nDesiredMatches)
        result[b0:b0 + rois.shape[0]] = picks.cpu().numpy()
    np.save(outputDir / "simsearch_indices_{}.npy".format(processTag), result, allow_pickle=True)
    return result


def main(outputDir, windowBins, blockSize, nCores, nDesiredMatches, nJobs, processTag):
    """similaritySearch_calc.py:13-34."""
    outputDir = Path(outputDir)
    print("Calculating search results...", flush=True)
    genomeCoords = np.load(outputDir / "genome_stats.npz", allow_pickle=True)["coords"]
    cube = np.load(outputDir / "simsearch_cube.npz", allow_pickle=True)
    roiCube, roiCoords = cube["scores"], cube["coords"]
    rowsToCalc = splitRows(roiCube.shape[0], nJobs)[processTag]
    return euclideanDistanceMulti(outputDir, genomeCoords, roiCoords, roiCube, windowBins, blockSize, nCores,
                                  nDesiredMatches, rowsToCalc, processTag)


if __name__ == "__main__":
    main(Path(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]),
         int(sys.argv[7]))
