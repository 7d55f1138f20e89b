"""Similarity-search preparation stage (similaritySearch_max_mean.py of the reference) and the score-text reader: the
host stage against fixtures produced by the unmodified reference (tests/golden/make_golden.py simsearch_prep), and the
oracle's restatement against the same fixtures.  Host code only: runs without a GPU."""
import gzip
import hashlib

import numpy as np
import pytest

from oracle import roi_oracle, simsearch_oracle as sso

CASES = ["simsearch_prep_real_chr1_60k", "simsearch_prep_synth_2chrom"]


def _digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def _write_scores(path, g):
    q, chrom, starts = g["scores_q"], g["chrom"], g["starts"]
    with gzip.open(path, "wt", compresslevel=1) as f:
        f.write("".join("%s\t%d\t%d\t%s\n" % (chrom[i], starts[i], starts[i] + 200, "\t".join("%.5f" % (v / 1e5) for v in q[i]))
                        for i in range(len(q))))


@pytest.mark.parametrize("name", CASES)
def test_prep_stage_matches_reference(tmp_path, golden, name):
    from epilogos_b200 import similaritySearch_max_mean as mm
    g = golden(name)
    path = tmp_path / "scores_x.txt.gz"
    _write_scores(path, g)
    for j in range(int(g["n_cfg"])):
        wb, bs, wbp, fst = (int(v) for v in g["cfg%d" % j])
        out = tmp_path / ("out%d" % j)
        out.mkdir()
        mm.main(out, path, wb, bs, wbp, fst, float(g["filter_score%d" % j]))
        stats = np.load(out / "genome_stats.npz", allow_pickle=True)
        assert np.array_equal(stats["scores"], g["scores_q"] / 1e5)
        assert stats["coords"].dtype == object and stats["coords"].shape == (len(g["scores_q"]), 3)
        assert list(stats["coords"][:, 0]) == list(g["chrom"]) and list(stats["coords"][:, 1]) == list(g["starts"])
        cube = np.load(out / "simsearch_cube.npz", allow_pickle=True)
        assert tuple(cube["scores"].shape) == tuple(g["cube_shape%d" % j])
        assert [str(c) for c in cube["coords"][:, 0]] == list(g["cube_chrom%d" % j])
        assert np.array_equal(cube["coords"][:, 1].astype(np.int64), g["cube_start%d" % j])
        assert np.array_equal(cube["coords"][:, 2].astype(np.int64), g["cube_end%d" % j])
        assert np.array_equal(cube["scores"][:3], g["cube_first%d" % j])
        assert np.array_equal(_digest(cube["scores"]), g["cube_digest%d" % j])
        red = np.load(out / "reduced_genome.npy", allow_pickle=True)
        assert len(red) == int(g["reduced_rows%d" % j])
        assert np.array_equal(_digest(red), g["reduced_digest%d" % j])


@pytest.mark.parametrize("name", CASES)
def test_prep_oracle_matches_reference(golden, name):
    """The CPU restatement (oracle/simsearch_oracle.py + oracle/roi_oracle.py) reproduces the reference's cube and
    reduced genome from the same scores."""
    g = golden(name)
    sc = g["scores_q"] / 1e5
    sums = sso.row_sums(sc)
    starts = g["starts"]
    for j in range(int(g["n_cfg"])):
        wb, bs, _, fst = (int(v) for v in g["cfg%d" % j])
        assert np.array_equal(_digest(sso.reduce_genome(sc, sums, bs)), g["reduced_digest%d" % j])
        sel = roi_oracle.max_mean(starts, starts + 200, sums, wb, len(sc) // wb)
        cube = np.stack([sso.make_slice(sc, sums, int(i), wb, bs) for i in sel["original_idx"]])
        coords = [(g["chrom"][i], int(s), int(e)) for i, s, e in zip(sel["original_idx"], sel["start"], sel["end"])]
        coords, cube = sso.remove_regions(coords, cube, fst, float(g["filter_score%d" % j]))
        assert [int(c[1]) for c in coords] == list(g["cube_start%d" % j])
        assert [str(c[0]) for c in coords] == list(g["cube_chrom%d" % j])
        assert np.array_equal(_digest(cube), g["cube_digest%d" % j])


def test_score_reader_fields_and_errors(tmp_path):
    from epilogos_b200 import helpers
    p = tmp_path / "s.txt"
    p.write_text("chr1\t0\t200\t0.00000\t-0.00000\t1.5\t12345.67891\n"
                 "\n"
                 "chr1\t200\t400\t1e-3\t-2.50000\t0.1\t123456789012345678.5\n"
                 "chrX\t0\t200\tnan\tinf\t-inf\t7")                      # unterminated last line is still a row (pandas)
    loc, sc = helpers.read_scores(p)
    assert sc.shape == (3, 4) and list(loc["chrom"]) == ["chr1", "chr1", "chrX"] and list(loc["end"]) == [200, 400, 200]
    assert sc[0, 0] == 0.0 and np.signbit(sc[0, 1]) and sc[0, 2] == 1.5 and sc[0, 3] == float("12345.67891")
    assert sc[1, 0] == 1e-3 and sc[1, 1] == -2.5 and sc[1, 2] == 0.1 and sc[1, 3] == float("123456789012345678.5")
    assert np.isnan(sc[2, 0]) and sc[2, 1] == np.inf and sc[2, 2] == -np.inf and sc[2, 3] == 7.0
    # every 5-decimal value the writer can emit converts to the nearest double (= Python's float())
    rng = np.random.default_rng(3)
    vals = np.round(rng.normal(0, 30, 4000), 5)
    q = tmp_path / "r.txt"
    q.write_text("".join("c\t%d\t%d\t%.5f\n" % (i, i + 1, v) for i, v in enumerate(vals)))
    _, got = helpers.read_scores(q)
    assert np.array_equal(got[:, 0], np.array([float("%.5f" % v) for v in vals]))
    bad = tmp_path / "bad.txt"
    bad.write_text("chr1\t0\t200\t0.1\t0.2\nchr1\t200\t400\t0.1\n")
    with pytest.raises(RuntimeError, match="score columns"):
        helpers.read_scores(bad)
    bad.write_text("chr1\t0\t200\t0.1\t0.2\nchr1\t200\t400\t0.1\tabc\n")
    with pytest.raises(RuntimeError, match="not a number"):
        helpers.read_scores(bad)
    with pytest.raises(FileNotFoundError):
        helpers.read_scores(tmp_path / "missing.txt")
