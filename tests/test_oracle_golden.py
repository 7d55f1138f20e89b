"""Pin the CPU oracle (oracle/epilogos_oracle.py) to outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which runs
expected.main -> expectedCombination.main -> scores.main of /root/reference.  Integer tables and the
float32 expected payload must match bit-for-bit; float32 scores must match bit-for-bit as well (the
oracle uses the same numpy ufuncs as the reference), and the formatted text byte-for-byte.
"""
import hashlib

import numpy as np
import pytest

from oracle import epilogos_oracle as orc

SINGLE = ["real10_chr1_k18", "real10_chr1_k18_nproc3", "synth_c833_k18", "synth_uniform_c833_k18", "synth_c127_k15"]


@pytest.mark.parametrize("name", SINGLE)
def test_s1_s2_tables_and_scores_match_reference(golden, name):
    g = golden(name)
    x, k = g["x"], int(g["num_states"])
    n1 = orc.s1_expected_counts(x, k)
    n2 = orc.s2_expected_counts(x, k)
    assert n1.dtype == g["s1_counts"].dtype and np.array_equal(n1, g["s1_counts"])
    assert n2.dtype == g["s2_counts"].dtype and np.array_equal(n2, g["s2_counts"])
    e1, e2 = orc.normalize_expected(n1), orc.normalize_expected(n2)
    assert e1.tobytes() == g["s1_exp"].tobytes()
    assert e2.tobytes() == g["s2_exp"].tobytes()
    assert orc.s1_scores(x, k, e1).tobytes() == g["s1_scores"].tobytes()
    assert orc.s2_scores(x, k, e2).tobytes() == g["s2_scores"].tobytes()


@pytest.mark.parametrize("name", ["real10_chr1_k18", "synth_c127_k15"])
def test_rowloop_port_equals_vectorised(golden, name):
    g = golden(name)
    x, k = g["x"][:300], int(g["num_states"])
    assert np.array_equal(orc.s1_expected_counts_rowloop(x, k), orc.s1_expected_counts(x, k))
    assert np.array_equal(orc.s2_expected_counts_rowloop(x, k), orc.s2_expected_counts(x, k))
    e1, e2 = g["s1_exp"], g["s2_exp"]
    assert orc.s1_scores_rowloop(x, k, e1).tobytes() == orc.s1_scores(x, k, e1).tobytes()
    assert orc.s2_scores_rowloop(x, k, e2).tobytes() == orc.s2_scores(x, k, e2).tobytes()


def test_scores_text_matches_reference_bytes(golden):
    g = golden("real10_chr1_k18")
    bins = g["x"].shape[0]
    # make_golden writes the slice with start_bin=0 (reference_driver.write_matrix_tsv)
    starts = np.arange(bins) * 200
    for s in (1, 2, 3):
        text = orc.format_scores_text(g["s%d_scores" % s], "chr1", starts, starts + 200)
        assert text == g["s%d_text" % s].tobytes()
        assert hashlib.sha256(text).digest() == g["s%d_text_sha256" % s].tobytes()


@pytest.mark.parametrize("name", ["real10_chr1_k18", "synth_s3_c12_k15", "synth_s3_c40_k18"])
def test_s3_tables_and_scores_match_reference(golden, name):
    g = golden(name)
    x, k = g["x"], int(g["num_states"])
    n3 = orc.s3_expected_counts(x, k)
    assert n3.dtype == g["s3_counts"].dtype and np.array_equal(n3, g["s3_counts"])
    c = x.shape[1]
    assert int(n3.sum()) == x.shape[0] * c * (c - 1)
    assert not n3[np.arange(c), np.arange(c)].any()
    assert np.array_equal(n3, n3.transpose(1, 0, 3, 2))
    e3 = orc.normalize_expected(n3)
    assert e3.tobytes() == g["s3_exp"].tobytes()
    sub = slice(0, 200)
    ref32 = orc.s3_scores_rowloop(x[sub], k, e3)
    assert ref32.tobytes() == g["s3_scores"][sub].tobytes()
    # the exact-arithmetic variant agrees with the reference's float32 to its accumulation noise
    f64 = orc.s3_scores_f64(x[sub], k, e3)
    assert np.max(np.abs(f64 - ref32)) < 2e-2
    assert np.array_equal(orc.s3_expected_counts_rowloop(x[:50], k), orc.s3_expected_counts(x[:50], k))


PAIRED = ["paired_real10_k18", "paired_synth_c30_c25_k18", "paired_synth_g20_k18", "paired_synth_q0_k18",
          "paired_synth_g40_k18"]


@pytest.mark.parametrize("name", PAIRED)
def test_paired_matches_reference(golden, name):
    g = golden(name)
    xa, xb, k = g["xa"], g["xb"], int(g["num_states"])
    seed, gs, q = int(g["seed"]), int(g["group_size"]), int(g["quiescent_state"])
    comb = np.concatenate((xa, xb), axis=1)
    perm = orc.reference_shuffle_indices(seed, comb.shape[0], comb.shape[1])
    bins = xa.shape[0]
    starts = np.arange(bins) * 200
    for s in (1, 2):
        if "s%d_counts" % s not in g.files:
            continue
        counts = orc.s1_expected_counts(comb, k) if s == 1 else orc.s2_expected_counts(comb, k)
        assert np.array_equal(counts, g["s%d_counts" % s])
        exp = orc.normalize_expected(counts)
        assert exp.tobytes() == g["s%d_exp" % s].tobytes()
        r = orc.paired_scores(xa, xb, perm, k, s, exp, q, gs)
        assert np.array_equal(r["quiescence"], g["s%d_quiescence" % s])
        assert r["null_distances"].dtype == np.float32
        assert r["null_distances"].tobytes() == g["s%d_null" % s].tobytes()
        text = orc.format_scores_text(r["delta"], "chr1", starts, starts + 200)
        assert text == g["s%d_delta_text" % s].tobytes()


def test_kl_masks():
    obs = np.array([0.0, 0.5, 0.25, 0.0])
    exp = np.array([0.5, 0.0, 0.25, 0.0], dtype=np.float32)
    out = orc.kl_terms(obs, exp)
    assert out.dtype == np.float64 and np.array_equal(out, np.zeros(4))
    assert orc.kl_terms(np.float32([0.5]), np.float32([0.25])).dtype == np.float32


def test_synth_generator_is_seeded():
    a = orc.synth_states(64, 33, 18, seed=5)
    b = orc.synth_states(64, 33, 18, seed=5)
    assert a.dtype == np.int8 and np.array_equal(a, b) and a.min() >= 0 and a.max() < 18


def test_full_real_chr1_matches_reference_digests(golden):
    """BASELINE configs[0] substitute: all 1,246,253 bins of the real chr1 matrix (10 biosamples, 18 states).  The
    reference's outputs are committed as digests (tests/golden/make_golden.py real_full_case)."""
    from oracle import roi_oracle
    g = golden("real10_chr1_full")
    x, k = g["x"], int(g["num_states"])
    n1, n2 = orc.s1_expected_counts(x, k), orc.s2_expected_counts(x, k)
    assert np.array_equal(n1, g["s1_counts"]) and np.array_equal(n2, g["s2_counts"])
    e1, e2 = orc.normalize_expected(n1), orc.normalize_expected(n2)
    assert e1.tobytes() == g["s1_exp"].tobytes() and e2.tobytes() == g["s2_exp"].tobytes()
    s1 = orc.s1_scores(x, k, e1)
    assert hashlib.sha256(np.ascontiguousarray(s1).tobytes()).digest() == g["s1_scores_sha256"].tobytes()
    s2 = orc.s2_scores(x, k, e2)
    assert hashlib.sha256(np.ascontiguousarray(s2).tobytes()).digest() == g["s2_scores_sha256"].tobytes()
    assert s2[:2000].tobytes() == g["s2_scores_head"].tobytes()
    starts = np.arange(len(s1), dtype=np.int64) * 200
    assert hashlib.sha256(orc.format_scores_text(s1, "chr1", starts, starts + 200)).digest() == \
        g["s1_text_sha256"].tobytes()
    sel = roi_oracle.max_mean(starts, starts + 200, s1.sum(axis=1), 50, 100)
    assert np.array_equal(sel["original_idx"], g["roi_original_idx"])          # the reference's own top-100 ranking
    assert np.array_equal(sel["start"], g["roi_start"]) and np.array_equal(sel["end"], g["roi_end"])
    assert sel["rolling_max"].tobytes() == g["roi_rolling_max"].tobytes()


def test_simsearch_oracle_matches_reference(golden):
    """SURVEY.md 8f row f4 (no product code yet): the restatement of similaritySearch_calc.runEuclideanDistance reproduces
    the unmodified reference's picks, including the threshold stop (-1 fill) of long result lists."""
    from oracle import simsearch_oracle as so
    g = golden("simsearch_g4000_k18")
    red = g["reduced_genome"]
    n = int(g["window_bins"]) // int(g["block_size"])
    for r, s in enumerate(g["roi_starts"]):
        assert np.array_equal(so.similar_regions(red, red[s:s + n], int(s), int(g["n_desired"])), g["indices"][r])
    for r, s in enumerate(g["roi_starts"][:2]):
        deep = so.similar_regions(red, red[s:s + n], int(s), g["indices_deep"].shape[1])
        assert np.array_equal(deep, g["indices_deep"][r]) and (deep == -1).any()
    g2 = golden("simsearch_g5000_k15_2chrom")                 # 15 states, 3-bin windows, every list ends by the threshold
    red2, n2 = g2["reduced_genome"], int(g2["window_bins"]) // int(g2["block_size"])
    for r, s in enumerate(g2["roi_starts"]):
        got = so.similar_regions(red2, red2[s:s + n2], int(s), int(g2["n_desired"]))
        assert np.array_equal(got, g2["indices"][r]) and (got == -1).any()
    d = so.window_distances(red, red[100:100 + n])
    # (the self-distance is ~1e-17, not 0: XX + YY - 2 X.Y^T rounds, which is why the reference excludes the ROI by position)
    assert d.shape == (len(red) - n + 1,) and np.argmin(d) == 100 and d[100] < 1e-12
    assert so.float_mode(np.array([3.0, 1.0, 3.0, 1.0, 2.0])) == 1.0
