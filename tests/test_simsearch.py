"""Similarity-search distance engine on the GPU against the unmodified reference's picks."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import simsearch_oracle as so            # noqa: E402  (checker only)


def _write_inputs(d, g, starts, n_super, block):
    red = g["reduced_genome"]
    nbins = len(red) * block
    coords = np.empty((nbins, 3), dtype=object)
    coords[:, 0] = "chr1"; coords[:, 1] = np.arange(nbins) * 200; coords[:, 2] = np.arange(nbins) * 200 + 200
    np.savez_compressed(d / "genome_stats", scores=np.zeros((1, 1)), coords=coords)
    roi_coords = np.empty((len(starts), 3), dtype=object)
    roi_coords[:, 0] = "chr1"
    roi_coords[:, 1] = [int(s) * block * 200 for s in starts]
    roi_coords[:, 2] = [(int(s) * block + n_super * block) * 200 for s in starts]
    np.savez_compressed(d / "simsearch_cube", scores=np.stack([red[s:s + n_super] for s in starts]), coords=roi_coords)
    np.save(d / "reduced_genome.npy", red)


def test_simsearch_matches_reference_golden(golden, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from epilogos_b200 import similaritySearch_calc as ssc
    g = golden("simsearch_g4000_k18")
    window, block = int(g["window_bins"]), int(g["block_size"])
    n_super = window // block
    starts = g["roi_starts"]
    _write_inputs(tmp_path, g, starts, n_super, block)
    got = ssc.main(tmp_path, window, block, 0, int(g["n_desired"]), 1, 0)
    assert got.dtype == np.int32 and np.array_equal(got, g["indices"])
    assert np.array_equal(np.load(tmp_path / "simsearch_indices_0.npy"), g["indices"])
    deep = ssc.main(tmp_path, window, block, 0, g["indices_deep"].shape[1], 3, 0)       # job 0 of 3 = the first ROI
    assert np.array_equal(deep, g["indices_deep"][:1]) and (deep == -1).any()
    # the device distances against the restatement of sklearn's formula, the mode against scipy's rule
    red = torch.from_numpy(g["reduced_genome"]).cuda()
    rois = torch.from_numpy(np.stack([g["reduced_genome"][s:s + n_super] for s in starts])).cuda()
    dist = ssc.window_distances(red, ssc.row_norms(red), rois).cpu().numpy()
    for r, s in enumerate(starts):
        ref = so.window_distances(g["reduced_genome"], g["reduced_genome"][s:s + n_super])
        np.testing.assert_allclose(dist[r], ref, rtol=1e-12, atol=1e-13)
    svals = torch.sort(torch.from_numpy(dist).cuda(), dim=1).values
    modes = ssc.mode_of_sorted(svals).cpu().numpy()
    for r in range(len(starts)):
        assert modes[r] == so.float_mode(dist[r])


def test_mode_of_sorted_edge_cases():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from epilogos_b200 import similaritySearch_calc as ssc
    rows = np.array([[1.0, 1.0, 2.0, 2.0, 3.0, 4.0, 5.0, 5.0],          # three runs of 2: the smallest value wins
                     [0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5],          # all distinct: the minimum
                     [2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0],          # one run
                     [0.0, 1.0, 1.0, 1.0, 7.0, 7.0, 7.0, 7.0]])         # the longer run at the end
    got = ssc.mode_of_sorted(torch.from_numpy(rows).cuda()).cpu().numpy()
    assert got.tolist() == [1.0, 0.5, 2.0, 7.0]
