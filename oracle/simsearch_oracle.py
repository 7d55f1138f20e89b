"""CPU restatement of the similarity-search distance engine -- TEST INFRASTRUCTURE ONLY.

Reference: similaritySearch_calc.runEuclideanDistance (similaritySearch_calc.py:67-123).  For every region of interest
(ROI, a window of nSuper = windowBins // blockSize reduced bins x K states) it computes the squared Euclidean distance to
every window of the reduced genome, takes half the MODE of those distances as the acceptance threshold, and greedily
picks up to nDesiredMatches non-overlapping windows in increasing distance (never the ROI itself).

No product code exists for this row yet; this restatement and its golden fixture
(tests/golden/simsearch_*.npz, produced by the unmodified reference) are the parity anchor for the kernel to come.

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


def similar_regions(reduced_genome, roi, region_start, n_desired):
    """similaritySearch_calc.py:101-123 for one ROI whose own window starts at reduced bin `region_start`.
    Returns int32 [n_desired] (unused slots keep 0 when the list ends by exhaustion, -1 after a threshold stop -- the
    reference pre-fills its output array with zeros and writes -1 only on the threshold branch)."""
    d = window_distances(reduced_genome, roi)
    n_super = np.asarray(roi).shape[0]
    half_mode = float_mode(d) / 2
    overlap = np.zeros(len(reduced_genome))
    overlap[region_start:region_start + n_super] = 1
    out = np.zeros(n_desired, dtype=np.int32)
    found = 0
    for hit in np.argsort(d):
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
    return out
