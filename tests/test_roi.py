"""Region-of-interest selection: oracle pinned to the reference's helpers.maxMean,
product (native epi_roi_maxmean + epilogos_b200.roi) against both."""
import numpy as np
import pytest

from oracle import epilogos_oracle as orc
from oracle import roi_oracle


def real_case(golden):
    g = golden("roi_real10_w50")
    x = g["x"]
    scores = orc.s1_scores(x, 18, g["exp"])
    n = len(scores)
    starts = np.arange(n, dtype=np.int64) * 200
    return g, scores, starts, starts + 200, np.array(["chr1"] * n, dtype=object)


def synth_case(golden, window):
    g = golden("roi_synth_w%d" % window)
    sc, st = g["scores"], g["starts"]
    n1 = int(g["chrom_split"])
    chrom = np.array(["chr1"] * n1 + ["chr2"] * (len(sc) - n1), dtype=object)
    return g, sc, st, st + 200, chrom


def check(sel, g):
    assert np.array_equal(sel["original_idx"], g["original_idx"])          # identical regions, identical ranking
    assert np.array_equal(sel["start"], g["start"]) and np.array_equal(sel["end"], g["end"])
    assert sel["rolling_max"].tobytes() == g["rolling_max"].tobytes()
    np.testing.assert_allclose(sel["rolling_mean"], g["rolling_mean"], rtol=1e-13, atol=1e-15)


def test_oracle_real_slice_matches_reference(golden):
    g, scores, st, en, _ = real_case(golden)
    check(roi_oracle.max_mean(st, en, scores.sum(axis=1), 50, 100), g)


@pytest.mark.parametrize("window", [50, 125, 7])
def test_oracle_two_chromosomes_matches_reference(golden, window):
    g, sc, st, en, _ = synth_case(golden, window)
    check(roi_oracle.max_mean(st, en, sc.sum(axis=1), window, 100), g)


def test_native_selector_matches_reference(golden):
    from epilogos_b200 import roi
    g, scores, st, en, _ = real_case(golden)
    check(roi.max_mean(st, en, scores.sum(axis=1), 50, 100), g)
    for window in (50, 125, 7):
        g, sc, st, en, _ = synth_case(golden, window)
        check(roi.max_mean(st, en, sc.sum(axis=1), window, 100), g)


def test_roi_file_and_max_states(golden, tmp_path):
    from epilogos_b200 import roi
    g, sc, st, en, chrom = synth_case(golden, 50)
    names = ["S%d" % i for i in range(1, 16)]
    sel = roi_oracle.max_mean(st, en, sc.sum(axis=1), 50, 100)
    states = roi_oracle.max_states(sc, sel["original_idx"], 50)
    want = roi_oracle.roi_text(chrom, sel, states, names)
    # product: temp_scores npz files per chromosome -> regionsOfInterest_<tag>.txt
    n1 = int(g["chrom_split"])
    for name, lo, hi in (("chr2", n1, len(sc)), ("chr1", 0, n1)):
        loc = np.empty((hi - lo, 3), dtype=object)
        loc[:, 0] = name; loc[:, 1] = st[lo:hi]; loc[:, 2] = en[lo:hi]
        np.savez_compressed(tmp_path / ("temp_scores_t_%s.npz" % name), chrName=np.array([name]), scoreArr=sc[lo:hi],
                            locationArr=loc)
    meta = tmp_path / "meta.tsv"
    meta.write_text("zero_index\tone_index\tshort_name\n" + "".join("%d\t%d\t%s\n" % (i, i + 1, n) for i, n in enumerate(names)))
    exp_path = tmp_path / "exp_freq_t.npy"
    np.save(exp_path, np.zeros(3, dtype=np.float32))
    roi.main(tmp_path, meta, "t", exp_path, 50, False)
    assert (tmp_path / "regionsOfInterest_t.txt").read_text() == want
    assert not exp_path.exists() and not list(tmp_path.glob("temp_scores_*.npz"))      # roiSingle.py:40, 73-74


def wide_case(golden):
    import hashlib
    g = golden("roi_wide_c833_k18")
    x = orc.synth_states(int(g["bins"]), int(g["cols"]), int(g["num_states"]), int(g["seed"]))
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest() == g["x_sha256"].tobytes(), \
        "the synthetic generator no longer reproduces the matrix the reference was run on"
    return g, x


@pytest.mark.parametrize("tag", ["s1", "s2"])
def test_wide_833_rankings_match_reference(golden, tag):
    """833 biosamples x 18 states, 24,000 bins: the oracle's float32 scores carry the reference's sha256, and both the
    oracle's and the native selector's top-100 equal the reference's own helpers.maxMean ranking, region by region."""
    import hashlib
    from epilogos_b200 import roi
    g, x = wide_case(golden)
    k = int(g["num_states"])
    scores = orc.s1_scores(x, k, g["s1_exp"]) if tag == "s1" else orc.s2_scores(x, k, g["s2_exp"])
    assert hashlib.sha256(np.ascontiguousarray(scores).tobytes()).digest() == g[tag + "_scores_sha256"].tobytes()
    starts = np.arange(len(x), dtype=np.int64) * 200
    for sel in (roi_oracle.max_mean(starts, starts + 200, scores.sum(axis=1), 50, 100),
                roi.max_mean(starts, starts + 200, scores.sum(axis=1), 50, 100)):
        assert np.array_equal(sel["original_idx"], g[tag + "_roi_original_idx"])
        assert np.array_equal(sel["start"], g[tag + "_roi_start"]) and np.array_equal(sel["end"], g[tag + "_roi_end"])
        assert sel["rolling_max"].tobytes() == g[tag + "_roi_rolling_max"].tobytes()
