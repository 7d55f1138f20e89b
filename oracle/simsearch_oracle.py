"""CPU restatement of the similarity-search distance engine -- TEST INFRASTRUCTURE ONLY.

Reference: similaritySearch_calc.runEuclideanDistance (similaritySearch_calc.py:67-123).  For every region of interest
(ROI, a window of nSuper = windowBins // blockSize reduced bins x K states) it computes the squared Euclidean distance to
every window of the reduced genome, takes half the MODE of those distances as the acceptance threshold, and greedily
picks up to nDesiredMatches non-overlapping windows in increasing distance (never the ROI itself).

Also restated: the preparation stage (similaritySearch_max_mean.py: window slices, region filters, genome reduction) and
the text of the output stage (similaritySearch_write.py).  The product code (csrc/simsearch.cu, csrc/hostio.cu and the
epilogos_b200.similaritySearch_* modules) never imports this file; the restatement and its fixtures
(tests/golden/simsearch_*.npz, produced by the unmodified reference through oracle/reference_driver.py) are the parity
anchor: checked against reference output in tests/test_oracle_golden.py and tests/test_simsearch_prep.py.

Parity notes found while pinning it:
  * the reference's distances come out of sklearn's euclidean_distances = XX + YY - 2 X.Y^T (a BLAS dgemm) clipped at 0,
    gathered along diagonals and summed over the window; restated here with the same numpy calls, bit for bit;
  * `st.mode` of floating-point distances is well defined only because identical windows give identical distances;
  * `np.argsort` is numpy's unstable introsort: the order among exactly tied distances is an implementation detail.
"""
import numpy as np


def window_distances(reduced_genome, roi):
    """similaritySearch_calc.py:86-99 -- sum over the window of squared distances between genome rows w+j and ROI rows j.
    reduced_genome: float [G, K]; roi: float [nSuper, K].  Returns float64 [G - nSuper + 1]."""
    x = np.asarray(reduced_genome)
    y = np.asarray(roi)
    n_super = y.shape[0]
    size = len(x) - (n_super - 1)
    # sklearn.metrics.pairwise.euclidean_distances(X, Y, squared=True) for float64 inputs
    xx = np.einsum("ij,ij->i", x, x)[:, np.newaxis]
    yy = np.einsum("ij,ij->i", y, y)[np.newaxis, :]
    d = -2 * np.dot(x, y.T)
    d += xx
    d += yy
    np.maximum(d, 0, out=d)
    idx0 = np.add(*np.broadcast_arrays(np.arange(n_super), np.arange(size).reshape(size, 1)))
    idx1 = np.broadcast_to(np.arange(n_super), (size, n_super))
    return np.sum(d[idx0, idx1], axis=1)


def float_mode(values):
    """scipy.stats.mode(values, keepdims=False)[0]: the most frequent value, the smallest one among ties."""
    u, c = np.unique(values, return_counts=True)
    return u[np.argmax(c)]


def similar_regions(reduced_genome, roi, region_start, n_desired, tie_order="numpy", return_distances=False):
    """similaritySearch_calc.py:101-123 for one ROI whose own window starts at reduced bin `region_start`.
    Returns int32 [n_desired] (unused slots keep 0 when the list ends by exhaustion, -1 after a threshold stop -- the
    reference pre-fills its output array with zeros and writes -1 only on the threshold branch).
    tie_order: windows at exactly equal distance are visited in the order of numpy's default (unstable, platform
    dependent) argsort, as the reference does ("numpy"), or in ascending window index ("index": what a stable sort gives
    and what the GPU engine does).  On real tracks adjacent windows of a repeated pattern tie exactly, so the two orders
    pick different members of a tie now and then; the distances of the picks are the same."""
    d = window_distances(reduced_genome, roi)
    n_super = np.asarray(roi).shape[0]
    half_mode = float_mode(d) / 2
    overlap = np.zeros(len(reduced_genome))
    overlap[region_start:region_start + n_super] = 1
    out = np.zeros(n_desired, dtype=np.int32)
    found = 0
    for hit in (np.argsort(d) if tie_order == "numpy" else np.argsort(d, kind="stable")):
        if np.any(overlap[hit:hit + n_super]):
            continue
        if d[hit] > half_mode:
            out[found:] = -1
            break
        out[found] = hit
        overlap[hit:hit + n_super] = 1
        found += 1
        if found >= n_desired:
            break
    return (out, d) if return_distances else out


# ------------------------------------------------------------------------------------------------
# preparation stage: similaritySearch_max_mean.py (window slices, filters, genome reduction)
# ------------------------------------------------------------------------------------------------

def row_sums(scores):
    """DataFrame.sum(axis=1) of the score frame (similaritySearch_max_mean.py:67, 97, 152): pandas adds the state
    columns one after another (checked against the reference: equal to the left-to-right sum, not to numpy's pairwise
    sum of a contiguous row)."""
    scores = np.asarray(scores, dtype=np.float64)
    out = np.zeros(len(scores), dtype=np.float64)
    for b in range(len(scores)):
        acc = 0.0
        for v in scores[b]:
            acc = acc + v
        out[b] = acc
    return out


def make_slice(scores, sums, idx, window_bins, block_size):
    """makeSlice (similaritySearch_max_mean.py:77-98): rows [idx - w//2, idx + w//2 (+1 if w odd)), grouped by position
    // blockSize, the first row with the largest sum of each group (groupby.idxmax)."""
    lo = idx - window_bins // 2
    hi = idx + window_bins // 2 + (1 if window_bins % 2 else 0)
    keep = []
    for g0 in range(lo, hi, block_size):
        best = g0
        for r in range(g0, min(g0 + block_size, hi)):
            if sums[r] > sums[best]:
                best = r
        keep.append(best)
    return np.asarray(scores)[keep]


def remove_regions(coords, cube, filter_state, filter_score):
    """removeRegions (similaritySearch_max_mean.py:101-134)."""
    keep = []
    for r in range(len(cube)):
        if int(coords[r][1]) >= int(coords[r][2]):
            continue
        if filter_state != 0:
            fs = cube.shape[2] - 1 if filter_state == -1 else filter_state - 1
            if int(np.argmax(cube[r].max(axis=0))) == fs:
                continue
        if filter_score != -1 and cube[r].max() < filter_score:
            continue
        keep.append(r)
    return [coords[r] for r in keep], cube[keep]


def reduce_genome(scores, sums, block_size):
    """reduceGenome (similaritySearch_max_mean.py:137-160): of every block of blockSize consecutive bins the one with the
    largest sum (sort by sum, drop_duplicates(keep='last')).  The reference's sort is unstable, so WHICH of several bins
    with exactly equal sums survives is not defined by the algorithm; this restatement takes the last one."""
    keep = []
    for g0 in range(0, len(scores), block_size):
        best = g0
        for r in range(g0, min(g0 + block_size, len(scores))):
            if sums[r] >= sums[best]:
                best = r
        keep.append(best)
    return np.asarray(scores)[keep]


# ------------------------------------------------------------------------------------------------
# output stage: similaritySearch_write.py (indices -> coordinates -> bed text)
# ------------------------------------------------------------------------------------------------

def bed_text(indices, genome_coords, roi_coords, window_bins, block_size):
    """The uncompressed text of simsearch.bed.gz (similaritySearch_write.py:44-67 reduced coordinates, :95-124 index ->
    coordinates, :142-152 one JSON list per region with the region itself first, sorted by (chrom, start))."""
    import json
    n = len(genome_coords)
    n_super = window_bins // block_size
    rows = []
    for r in range(len(indices)):
        recs = ["%s:%s:%s" % (roi_coords[r][0], roi_coords[r][1], roi_coords[r][2])]
        for w in indices[r]:
            if w == -1:
                continue
            first = int(w) * block_size
            last = min((int(w) + n_super - 1) * block_size + block_size - 1, n - 1)
            recs.append("%s:%s:%s" % (genome_coords[first][0], genome_coords[first][1], genome_coords[last][2]))
        rows.append((roi_coords[r][0], roi_coords[r][1], roi_coords[r][2], json.dumps(recs)))
    rows.sort(key=lambda t: (t[0], t[1]))
    return "".join("%s\t%s\t%s\t%s\n" % t for t in rows).encode()
