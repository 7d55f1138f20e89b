"""Paired-mode score stage.  Mirror of scores.calculateScoresPairwise (scores.py:172-256) and of the paired
branches of s1Score / s2Score (scores.py:282-303, 319-322, 373-398, 414-421).

Outputs (rank 0): `pairwiseDelta_<tag>_<file>.txt.gz`, `temp_nullDistances_<tag>_<file>.npz`
{chrName, nullDistances float32[B]}, `temp_quiescence_<tag>_<file>.npz` {chrName, quiescenceArr bool[B]}.

Null shuffles.  The reference shuffles each combined row once with the unseeded global numpy RNG
(helpers.py:183-184).  Two modes are provided (environment variable EPILOGOS_B200_NULL or the `null_mode`
argument):
  "reference" -- draw argsort(np.random.rand(rows, N)) on the host exactly as the reference does and replay
                 those indices on the GPU: bit-reproducible against a seeded single-process reference run.
  "device"    -- (default) draw the shuffle on the GPU from the per-bin counts with a counter-based Philox
                 stream; same distribution, no host RNG, supports `nperm` > 1 shuffles per bin.
"""
import os
from pathlib import Path

import numpy as np

from . import dist, helpers, session, writer
from .scores import _gather_locations


def calculateScoresPairwise(saliency, file1Path, file2Path, numStates, outputDirPath, expFreqPath, fileTag, filename,
                            quiescentState, groupSize, verbose, backend=None, null_mode=None, seed=0, nperm=1):
    if saliency not in (1, 2):
        raise ValueError("Please ensure that saliency metric is either 1 or 2 for Pairwise Epilogos")
    null_mode = null_mode or os.environ.get("EPILOGOS_B200_NULL", "device")
    if null_mode not in ("reference", "device"):
        raise ValueError("null_mode must be 'reference' or 'device'")
    be = session.get_backend(backend)
    shard = session.load_shard(file1Path, file2Path, numStates, backend)
    c1, c2 = shard.states_a.shape[1], shard.states_b.shape[1]
    size_a, size_b = (c1, c2) if groupSize == -1 else (groupSize, groupSize)       # helpers.py:190-194
    exp = be.to_device(np.load(expFreqPath, allow_pickle=False))

    cnt_a, cnt_b = shard.counts("a"), shard.counts("b")
    if null_mode == "reference":
        rows = shard.states_a.shape[0]
        perm = np.argsort(np.random.rand(rows, c1 + c2), axis=1)                     # helpers.py:183
        null_a, null_b = be.shuffled_counts_perm(shard.states_a, shard.states_b, perm, numStates, size_a, size_b)
    else:
        null_a, null_b = be.shuffled_counts_device(cnt_a, cnt_b, size_a, size_b, seed + 7919 * dist.rank(), nperm,
                                                   width=c1 + c2)

    # S1 observes over the width of the array it is given (scores.py:343); S2 normalises the shuffled halves with
    # the ORIGINAL group widths even under -g (scores.py:397-398, 418-421)
    p1, p2 = c1 * (c1 - 1), c2 * (c2 - 1)
    score_a = be.scores(cnt_a, c1, saliency, exp, perms=p1)
    score_b = be.scores(cnt_b, c2, saliency, exp, perms=p2)
    nshape = tuple(null_a.shape)
    flat_a, flat_b = null_a.reshape(-1, numStates), null_b.reshape(-1, numStates)
    nscore_a = be.scores(flat_a, size_a if saliency == 1 else max(size_a, 1), saliency, exp, perms=p1)
    nscore_b = be.scores(flat_b, size_b if saliency == 1 else max(size_b, 1), saliency, exp, perms=p2)
    delta, _ = be.pairwise_combine(score_a, score_b, None, None)
    _, null_dist = be.pairwise_combine(None, None, nscore_a, nscore_b)
    if len(nshape) == 3:
        null_dist = null_dist.reshape(nshape[0], nshape[1]).transpose(0, 1).contiguous().reshape(nshape[0], nshape[1])
    quies = be.quiescent_mask(cnt_a, c1, cnt_b, c2, quiescentState)

    total = shard.total_rows
    delta = dist.gather_rows(delta, total)
    quies = dist.gather_rows(quies, total)
    if null_dist.dim() == 2:            # [nperm, rows] -> gather per permutation, keep permutation-major order
        parts = [dist.gather_rows(null_dist[p].contiguous(), total) for p in range(null_dist.shape[0])]
        null_dist = None if parts[0] is None else __import__("torch").cat(parts, dim=0)
    else:
        null_dist = dist.gather_rows(null_dist, total)
    loc = _gather_locations(shard)
    if dist.rank() == 0:
        outputDirPath = Path(outputDirPath)
        writer.write_scores_text(outputDirPath / "pairwiseDelta_{}_{}.txt.gz".format(fileTag, filename),
                                 delta.cpu().numpy(), loc)
        chrName = loc["chrom"][0] if len(loc["chrom"]) else ""
        helpers.savez_level(outputDirPath / "temp_nullDistances_{}_{}.npz".format(fileTag, filename),
                            chrName=np.array([chrName]), nullDistances=null_dist.cpu().numpy())
        helpers.savez_level(outputDirPath / "temp_quiescence_{}_{}.npz".format(fileTag, filename),
                            chrName=np.array([chrName]), quiescenceArr=quies.cpu().numpy().astype(np.bool_))
