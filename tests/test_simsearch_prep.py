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


def _prepared(tmp_path, golden):
    from epilogos_b200 import similaritySearch_max_mean as mm
    g = golden("simsearch_prep_real_chr1_60k")
    path = tmp_path / "scores_x.txt.gz"
    _write_scores(path, g)
    out = tmp_path / "build"
    out.mkdir()
    mm.main(out, path, 125, 5, 25000, -1, -1.0)
    return path, out


def test_write_stage_matches_reference_text(tmp_path, golden):
    """similaritySearch_write.main on the reference's own index array (split over three job files, as its SLURM jobs
    leave them): same bed text, same files left behind, and a well-formed BGZF container."""
    import struct
    from epilogos_b200 import similaritySearch_write as ssw
    from epilogos_b200.helpers import splitRows
    c = golden("simsearch_chain_real_chr1_60k")
    _, out = _prepared(tmp_path, golden)
    idx = c["indices"]
    for j, (lo, hi) in enumerate(splitRows(len(idx), 3)):
        np.save(out / ("simsearch_indices_%d.npy" % j), idx[lo:hi])
    ssw.main(out, 125, 5, 3, idx.shape[1])
    with gzip.open(out / "simsearch.bed.gz", "rb") as f:
        text = f.read()
    assert len(text) == int(c["bed_bytes"]) and text[:2000] == c["bed_head"].tobytes()
    assert np.array_equal(np.frombuffer(hashlib.sha256(text).digest(), dtype=np.uint8), c["bed_digest"])
    assert np.array_equal(np.load(out / "simsearch_indices.npy"), idx)
    left = sorted(p.name for p in out.iterdir())
    assert left == [n for n in c["leftovers"] if not n.endswith(".tbi")]           # the tabix index is not produced here
    # BGZF: every member carries the BC extra field with its own size, the file ends with the empty EOF member
    raw = (out / "simsearch.bed.gz").read_bytes()
    off, blocks = 0, 0
    while off < len(raw):
        assert raw[off:off + 4] == b"\x1f\x8b\x08\x04" and raw[off + 12:off + 16] == b"BC\x02\x00"
        off += struct.unpack_from("<H", raw, off + 16)[0] + 1
        blocks += 1
    assert off == len(raw) and blocks >= 2 and raw[-28:] == ssw._BGZF_EOF
    # the oracle's restatement of the text from the same inputs
    cube = np.load(out / "simsearch_cube.npz", allow_pickle=True)
    g = golden("simsearch_prep_real_chr1_60k")
    coords = [(ch, int(s), int(s) + 200) for ch, s in zip(g["chrom"], g["starts"])]
    assert sso.bed_text(idx, coords, cube["coords"], 125, 5) == text


def test_query_mode_and_cli_validation(tmp_path, golden):
    from click.testing import CliRunner
    from epilogos_b200 import similaritySearch_run as ssr, similaritySearch_write as ssw
    lines = ('chr1\t1000\t26000\t["chr1:1000:26000", "chr2:5000:30000", "chrX:0:25000"]\n'
             'chr1\t50000\t75000\t["chr1:50000:75000"]\n'
             'chr2\t0\t25000\t["chr2:0:25000", "chr1:1000:26000"]\n')
    bed = tmp_path / "simsearch.bed.gz"
    ssw.bgzf_write(bed, lines.encode())
    qdir = tmp_path / "q"
    res = CliRunner().invoke(ssr.main, ["-q", "chr1:0-30000", "-m", str(bed), "-o", str(qdir)])
    assert res.exit_code == 0, res.output
    assert (qdir / "similarity_search_region_chr1_1000_26000_recs.bed").read_text() == "chr2\t5000\t30000\nchrX\t0\t25000\n"
    qfile = tmp_path / "queries.bed"
    qfile.write_text("chr1\t40000\t80000\nchr2\t0\t26000\nchr9\t0\t10\n")
    written = ssr.querySimSearch(str(qfile), bed, qdir)
    assert [w.name for w in written] == ["similarity_search_region_chr1_50000_75000_recs.bed",
                                         "similarity_search_region_chr2_0_25000_recs.bed"]
    assert written[0].read_text() == "" and written[1].read_text() == "chr1\t1000\t26000\n"
    with pytest.raises(ValueError, match="valid query"):
        ssr.generateRegionArr("1:5-9")
    for args, msg in ((["-o", str(qdir)], "Either -b or -q"), (["-b", "-q", "chr1:1-2", "-o", str(qdir)], "Both -b and -q")):
        res = CliRunner().invoke(ssr.main, args)
        assert isinstance(res.exception, ValueError) and msg in str(res.exception)
    assert ssr.determineBlockSize200(25000) == 5 and ssr.determineBlockSize20(10000) == 20
    with pytest.raises(ValueError, match="window size"):
        ssr.determineBlockSize200(30000)
    p = tmp_path / "s.txt"
    p.write_text("chr1\t400\t600\t0.1\n")
    assert ssr.determineBinSize(p) == 200


def test_build_driver_host_chain_with_replayed_engine(tmp_path, golden, monkeypatch):
    """similaritySearch_run.buildSimSearch wiring on CPU: bin size -> window / block size, preparation, per-job index
    files, writer, clean-up.  The GPU distance engine (the one stage that needs a device) is replaced by a stand-in that
    replays the reference's picks for the regions it is asked for, so everything else must reproduce the reference's
    index array and bed text bit for bit."""
    from epilogos_b200 import similaritySearch_calc, similaritySearch_run as ssr
    from epilogos_b200.helpers import splitRows
    prep = golden("simsearch_prep_real_chr1_60k")
    c = golden("simsearch_chain_real_chr1_60k")
    path = tmp_path / "scores_x.txt.gz"
    _write_scores(path, prep)
    out = tmp_path / "build"
    calls = []

    def replay_engine(outputDir, windowBins, blockSize, nCores, nDesiredMatches, nJobs, processTag):
        calls.append((windowBins, blockSize, nDesiredMatches, nJobs, processTag))
        cube = np.load(outputDir / "simsearch_cube.npz", allow_pickle=True)
        assert len(cube["scores"]) == len(c["indices"]) and (outputDir / "reduced_genome.npy").exists()
        lo, hi = splitRows(len(cube["scores"]), nJobs)[processTag]
        np.save(outputDir / "simsearch_indices_{}.npy".format(processTag), c["indices"][lo:hi])

    monkeypatch.setattr(similaritySearch_calc, "main", replay_engine)
    out.mkdir()
    idx = ssr.buildSimSearch(path, out, -1, 100, -1, -1.0)
    assert calls == [(125, 5, 100, 1, 0)]
    assert np.array_equal(idx, c["indices"])
    with gzip.open(out / "simsearch.bed.gz", "rb") as f:
        text = f.read()
    assert np.array_equal(np.frombuffer(hashlib.sha256(text).digest(), dtype=np.uint8), c["bed_digest"])
    assert sorted(p.name for p in out.iterdir()) == [n for n in c["leftovers"] if not n.endswith(".tbi")]
    bad = tmp_path / "bins50.txt"
    bad.write_text("chr1\t0\t50\t0.1\n")
    with pytest.raises(ValueError, match="200bp or 20bp"):
        ssr.buildSimSearch(bad, out, -1, 100, -1, -1.0)


def test_score_reader_decimal_forms_equal_python_float(tmp_path):
    """Every way a float can be written in a score file converts to the same double as Python's float(): fixed
    notation with 0..20 decimals (the fast path up to 15 significant digits, strtod beyond), exponents, repr(), big
    integers, leading zeros, an explicit plus sign."""
    from epilogos_b200 import helpers
    rng = np.random.default_rng(11)
    texts = []
    for v in np.concatenate((rng.normal(0, 1, 300), rng.normal(0, 1e4, 300), rng.uniform(-1e-6, 1e-6, 200),
                             rng.integers(-10**15, 10**15, 100).astype(np.float64))):
        v = float(v)
        texts += ["%.*f" % (int(rng.integers(0, 21)), v), "%e" % v, "%.17g" % v, repr(v), "%.5f" % v]
    texts += ["0", "-0", "+1.5", "000123.4500", "1.", ".5", "123456789012345", "1234567890123456", "9007199254740993",
              "0.1234567890123456789", "12345678901234567890123", "1e22", "1E-7", "4.9e-324", "1.7976931348623157e308"]
    p = tmp_path / "forms.txt"
    p.write_text("".join("c\t%d\t%d\t%s\n" % (i, i + 1, t) for i, t in enumerate(texts)))
    _, got = helpers.read_scores(p)
    want = np.array([float(t) for t in texts])
    bad = [(t, g, w) for t, g, w in zip(texts, got[:, 0], want) if not (g == w and np.signbit(g) == np.signbit(w))]
    assert not bad, bad[:5]


def test_tie_order_changes_only_exactly_tied_picks(tmp_path, golden):
    """The parity caveat of the distance engine, pinned on CPU: on the real-data fixture the reference's picks (numpy's
    unstable argsort) and the picks with ties visited in ascending window index (a stable sort, what the GPU engine does)
    differ in 5 of 358 regions, and wherever they differ the two windows are at exactly the same distance."""
    _, out = _prepared(tmp_path, golden)
    prep = golden("simsearch_prep_real_chr1_60k")
    ref = golden("simsearch_chain_real_chr1_60k")["indices"]
    cube = np.load(out / "simsearch_cube.npz", allow_pickle=True)
    red = np.load(out / "reduced_genome.npy")
    known = [282, 336, 337, 343, 356]
    for r in known + [0, 50, 120, 200, 300]:
        s0 = int(np.flatnonzero(prep["starts"] == cube["coords"][r][1])[0]) // 5
        by_index, d = sso.similar_regions(red, cube["scores"][r], s0, ref.shape[1], tie_order="index", return_distances=True)
        assert np.array_equal(by_index == -1, ref[r] == -1)
        keep = ref[r] != -1
        assert np.array_equal(d[by_index[keep]], d[ref[r][keep]])
        assert np.array_equal(by_index, ref[r]) == (r not in known)
