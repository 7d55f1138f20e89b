"""Native host I/O (epi_pack_tsv / epi_write_scores_gz / epi_tsv_shape): no GPU needed."""
import gzip
import sys
from pathlib import Path

import numpy as np
import pytest

from epilogos_b200 import helpers, writer
from epilogos_b200._lib import EpilogosB200Error

ROOT = Path(__file__).resolve().parent.parent


def test_writer_matches_python_format_on_edge_values(tmp_path):
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        np.float32([0.0, -0.0, 1e-7, -1e-7, 5e-6, -5e-6, 1.0 / 64, 3.0 / 64, -5.0 / 64, 0.000005, 0.000015, 0.000025,
                    123456.789, -98765.4321, 1e9, -3e12, 1e-30, 0.5, 0.99999, 0.999995, 0.9999949]),
        (rng.standard_normal(20000) * rng.choice([1e-6, 1e-3, 1.0, 1e3], 20000)).astype(np.float32),
        (rng.integers(-10 ** 6, 10 ** 6, 20000) / 64.0).astype(np.float32),          # many exact decimal ties
        (rng.integers(-10 ** 7, 10 ** 7, 20000) * 1e-5).astype(np.float32),
    ])
    k = 7
    vals = vals[: (len(vals) // k) * k].reshape(-1, k)
    n = len(vals)
    loc = dict(chrom=np.array(["chrX"] * (n // 2) + ["chr2_random"] * (n - n // 2), dtype=object),
               start=np.arange(n, dtype=np.int64) * 200, end=np.arange(n, dtype=np.int64) * 200 + 200)
    path = tmp_path / "s.txt.gz"
    writer.write_scores_text(path, vals, loc, threads=3)
    got = gzip.open(path, "rb").read().decode().splitlines()
    assert len(got) == n
    for i in (list(range(40)) + list(rng.integers(0, n, 3000))):
        want = "%s\t%d\t%d\t%s" % (loc["chrom"][i], loc["start"][i], loc["end"][i],
                                  "\t".join("{:.5f}".format(v) for v in vals[i]))
        assert got[i] == want, (i, got[i], want)
    # whole-file check against Python formatting
    ref = "".join("%s\t%d\t%d\t%s\n" % (loc["chrom"][i], loc["start"][i], loc["end"][i],
                                       "\t".join("%.5f" % float(v) for v in vals[i])) for i in range(n))
    assert "\n".join(got) + "\n" == ref


def test_writer_empty(tmp_path):
    path = tmp_path / "e.txt.gz"
    loc = dict(chrom=np.array([], dtype=object), start=np.zeros(0, np.int64), end=np.zeros(0, np.int64))
    writer.write_scores_text(path, np.zeros((0, 18), np.float32), loc)
    assert gzip.open(path, "rb").read() == b""


def test_packer_shapes_ranges_and_validation(tmp_path):
    f = tmp_path / "m.txt"
    rows = ["chr1\t%d\t%d\t1\t18\t7" % (i * 200, i * 200 + 200) for i in range(10)]
    f.write_text("\n".join(rows) + "\n")
    assert helpers.tsv_shape(f) == (10, 3) and helpers.countRows(f) == 10
    loc, s0 = helpers.read_matrix(f, num_states=18)
    assert s0.tolist() == [[0, 17, 6]] * 10 and s0.base.shape[1] == 16 and not s0.base[:, 3:].any()
    _, part = helpers.read_matrix(f, rows=(3, 7), num_states=18, want_locations=False)
    assert part.shape == (4, 3)
    # no trailing newline: the reference counts newline characters, so the last line is not a row
    g = tmp_path / "n.txt"
    g.write_text("\n".join(rows))
    assert helpers.countRows(g) == 9
    # CRLF line ends
    h = tmp_path / "c.txt"
    h.write_bytes(("\r\n".join(rows) + "\r\n").encode())
    assert helpers.read_matrix(h, num_states=18)[1].tolist() == [[0, 17, 6]] * 10
    for bad, msg in [("chr1\t0\t200\t1\t19\t7\n", "outside 1..18"), ("chr1\t0\t200\t1\t0\t7\n", "outside 1..18"),
                     ("chr1\t0\t200\t1\tx\t7\n", "not an integer"), ("chr1\t0\t200\t1\t2\n", "state columns"),
                     ("chr1\t0\t200\t1\t2\t3\t4\n", "more than 3")]:
        b = tmp_path / "bad.txt"
        b.write_text(rows[0] + "\n" + bad)
        with pytest.raises(EpilogosB200Error, match=msg):
            helpers.read_matrix(b, num_states=18)
    with pytest.raises(FileNotFoundError):
        helpers.read_matrix(tmp_path / "missing.txt")
    with pytest.raises(EpilogosB200Error, match="outside the 10 parsed rows"):
        helpers.read_matrix(f, rows=(20, 30), num_states=18)                       # one-pass parser
    with pytest.raises(EpilogosB200Error, match="only"):
        helpers.read_matrix(f, rows=(20, 30), num_states=18, shape=(10, 3))       # two-call packer


def test_one_pass_parser_equals_two_call_packer(tmp_path):
    """epi_tsv_parse_* (row count not known beforehand, one inflate pass) against epi_tsv_shape + epi_pack_tsv: same labels,
    pad bytes, coordinates and chromosome table for whole files, row ranges and rank splits; an unterminated last line is
    not a row in either (the reference counts newline characters); an empty file has no rows."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 18, size=(70001, 37)).astype(np.int8)                  # spans three internal 32768-row chunks
    f = tmp_path / "m.txt.gz"
    with gzip.open(f, "wt", compresslevel=1) as out:
        for r in range(len(x)):
            out.write("chr%d\t%d\t%d\t%s\n" % (1 + r // 40000, r * 200, r * 200 + 200, "\t".join(map(str, (x[r] + 1).tolist()))))
        out.write("chr2\t1\t2\t" + "\t".join(["1"] * 37))                        # no newline: not a row
    shape = helpers.tsv_shape(f)
    assert shape == (70001, 37)
    for rows in (None, (0, 70001), (32767, 32769), (65536, 70001), (500, 500)):
        la, a, total = helpers.read_matrix(f, rows, num_states=18, return_total=True)
        lb, b = helpers.read_matrix(f, rows if rows is not None else (0, 70001), num_states=18, shape=shape)
        assert total == 70001 and np.array_equal(a, b) and np.array_equal(a.base, b.base)
        for key in ("start", "end"):
            assert np.array_equal(la[key], lb[key])
        assert list(la["chrom"]) == list(lb["chrom"])          # (ids number the file's names in one, the range's in the other)
        names = la["chrom_names"].split(b"\0")
        assert [names[i].decode() for i in la["chrom_id"]] == list(la["chrom"])
    lo, hi = helpers.splitRows(70001, 3)[1]
    _, part = helpers.read_matrix(f, num_states=18, split=(1, 3), want_locations=False)
    assert np.array_equal(part, x[lo:hi])
    e = tmp_path / "empty.txt"
    e.write_text("")
    loc, s0, total = helpers.read_matrix(e, num_states=18, return_total=True)
    assert total == 0 and s0.shape[0] == 0


def test_packer_long_lines_across_buffer_blocks(tmp_path):
    """Rows longer than any internal block boundary handling: 40 000 rows x 700 columns, gzip."""
    rng = np.random.default_rng(1)
    x = rng.integers(0, 25, size=(40000, 700)).astype(np.int8)
    f = tmp_path / "big.txt.gz"
    with gzip.open(f, "wt", compresslevel=1) as out:
        for r in range(len(x)):
            out.write("chr%d\t%d\t%d\t%s\n" % (1 + r // 30000, r * 200, r * 200 + 200, "\t".join(map(str, (x[r] + 1).tolist()))))
    loc, s0 = helpers.read_matrix(f, num_states=25)
    assert np.array_equal(s0, x)
    assert loc["chrom"][0] == "chr1" and loc["chrom"][-1] == "chr2" and loc["start"][-1] == 39999 * 200


# ---- the reader's own DEFLATE / gzip decoder (csrc/fast_inflate.h) against zlib ------------------------------------
def _inflate(path, monkeypatch=None, use_zlib=False):
    import ctypes
    from epilogos_b200 import _lib
    if monkeypatch is not None:
        if use_zlib:
            monkeypatch.setenv("EPI_ZLIB_INFLATE", "1")
        else:
            monkeypatch.delenv("EPI_ZLIB_INFLATE", raising=False)
    n = ctypes.c_int64(0)
    _lib.call("epi_inflate_file", str(path).encode(), ctypes.c_void_p(0), 0, ctypes.byref(n))
    buf = np.empty(max(n.value, 1), dtype=np.uint8)
    _lib.call("epi_inflate_file", str(path).encode(), ctypes.c_void_p(buf.ctypes.data), n.value, ctypes.byref(n))
    return buf[:n.value].tobytes()


def test_native_inflate_equals_zlib_on_every_block_type(tmp_path, monkeypatch):
    """Stored, fixed-Huffman and dynamic-Huffman blocks, long codes (random bytes at level 9), short-period matches,
    literal-only streams, sync / full flushes (empty stored blocks), multi-member files (what the score writer emits),
    empty input, and a 40 MB stream that crosses the reader's 16 MB blocks in the middle of matches."""
    import zlib
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"0123456789\t\n", dtype=np.uint8)
    row = b"chr1\t0\t200\t" + b"\t".join(str(int(v)).encode() for v in rng.integers(1, 19, 833)) + b"\n"
    cases = {
        "empty": b"", "one": b"x", "text": b"chr1\t0\t200\t1\t2\t3\n" * 1000,
        "random": rng.integers(0, 256, 300000, dtype=np.uint8).tobytes(),
        "lowentropy": rng.choice(alphabet, 2_000_000, p=[.3, .2, .1, .05, .05, .05, .05, .04, .03, .03, .08, .02]).tobytes(),
        "runs": b"a" * 70000 + b"ab" * 40000 + b"abc" * 30000 + b"abcdefg" * 20000 + bytes(range(256)) * 3000,
        "blocks40MB": (row * 7 + rng.integers(48, 58, 1000, dtype=np.uint8).tobytes()) * 2000,
    }
    for name, data in cases.items():
        variants = {}
        for lvl in (0, 1, 6, 9) if len(data) < 10_000_000 else (1,):
            variants["l%d" % lvl] = gzip.compress(data, lvl)
        if len(data) < 10_000_000:
            for tag, strategy in (("fixed", zlib.Z_FIXED), ("huffman", zlib.Z_HUFFMAN_ONLY), ("rle", zlib.Z_RLE)):
                co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, strategy)
                variants[tag] = co.compress(data) + co.flush()
            step = max(len(data) // 7, 1)
            variants["multi"] = b"".join(gzip.compress(data[i:i + step], 1 + (i % 9)) for i in range(0, max(len(data), 1), step))
            co = zlib.compressobj(6, zlib.DEFLATED, 31)
            out = b""
            for j, i in enumerate(range(0, len(data), 50021)):
                out += co.compress(data[i:i + 50021]) + co.flush(zlib.Z_SYNC_FLUSH if j % 2 else zlib.Z_FULL_FLUSH)
            variants["flush"] = out + co.flush()
            variants["padded"] = gzip.compress(data, 6) + b"\0" * 37            # zero padding after the last member
        for tag, packed in variants.items():
            p = tmp_path / ("%s_%s.gz" % (name, tag))
            p.write_bytes(packed)
            assert _inflate(p, monkeypatch) == data, (name, tag)
    p = tmp_path / "plain.txt"
    p.write_bytes(cases["text"])
    assert _inflate(p, monkeypatch) == cases["text"]                            # not gzip: passed through
    p = tmp_path / "text_l6.gz"
    assert _inflate(p, monkeypatch, use_zlib=True) == cases["text"]             # EPI_ZLIB_INFLATE=1: the zlib path


def test_damaged_gzip_is_an_error_not_a_short_read(tmp_path, monkeypatch):
    """A stream that is cut short, fails its CRC-32 or holds invalid codes raises (the pandas reader of the reference
    raises too); it is never returned as a shorter matrix.  Both decoders."""
    rng = np.random.default_rng(1)
    text = b"".join(b"chr1\t%d\t%d\t" % (i * 200, i * 200 + 200) + b"\t".join(str(int(v)).encode() for v in rng.integers(1, 19, 40)) + b"\n"
                    for i in range(5000))
    good = gzip.compress(text, 6)
    p = tmp_path / "m.txt.gz"
    damaged = {"cut in the data": good[:len(good) // 2], "cut in the trailer": good[:-3],
               "crc": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:], "length": good[:-1] + bytes([good[-1] ^ 0x40]),
               "data": good[:len(good) // 3] + bytes([good[len(good) // 3] ^ 0x10]) + good[len(good) // 3 + 1:]}
    for use_zlib in (False, True):
        for what, blob in damaged.items():
            p.write_bytes(blob)
            with pytest.raises(EpilogosB200Error):
                _inflate(p, monkeypatch, use_zlib)
            with pytest.raises(EpilogosB200Error):
                helpers.read_matrix(p, num_states=18)
        p.write_bytes(good)
        loc, states = helpers.read_matrix(p, num_states=18)
        assert states.shape == (5000, 40) and _inflate(p, monkeypatch, use_zlib) == text
    # seeded single-bit damage anywhere in the file: rejected, or (header fields that carry no data) decoded in full
    for trial in range(300):
        blob = bytearray(good)
        i = int(rng.integers(2, len(blob)))
        blob[i] ^= 1 << int(rng.integers(0, 8))
        p.write_bytes(bytes(blob))
        try:
            got = _inflate(p, monkeypatch)
        except EpilogosB200Error:
            continue
        assert got == text, "byte %d flipped: accepted with different content" % i


def test_parser_fast_and_careful_paths_agree(tmp_path):
    """One-, two- and three-digit labels mixed in every column position (the fast path takes the first two shapes, the
    careful loop the rest and the last column), single-column files, and errors raised from inside long rows."""
    rng = np.random.default_rng(5)
    for cols, k in ((1, 5), (2, 9), (3, 127), (40, 127), (833, 18), (200, 100)):
        x = rng.integers(1, k + 1, (300, cols))
        p = tmp_path / ("m_%d_%d.txt" % (cols, k))
        p.write_text("".join("chrZ\t%d\t%d\t%s\n" % (i * 200, i * 200 + 200, "\t".join(map(str, r))) for i, r in enumerate(x.tolist())))
        loc, got = helpers.read_matrix(p, num_states=k)
        assert np.array_equal(got, x - 1) and list(loc["start"][:2]) == [0, 200][:len(x)]
    x = rng.integers(1, 19, (50, 600))
    rows = ["chr1\t%d\t%d\t%s" % (i * 200, i * 200 + 200, "\t".join(map(str, r))) for i, r in enumerate(x.tolist())]
    for bad_field, msg in (("19", "outside 1..18"), ("0", "outside 1..18"), ("1x", "unexpected character"), ("", "not an integer"),
                           ("-3", "not an integer")):
        broken = list(rows)
        f = broken[17].split("\t")
        f[3 + 411] = bad_field
        broken[17] = "\t".join(f)
        p = tmp_path / "bad.txt"
        p.write_text("\n".join(broken) + "\n")
        with pytest.raises(EpilogosB200Error, match="row 17.*" + msg):
            helpers.read_matrix(p, num_states=18)


def test_reader_thread_budget_divides_the_host(monkeypatch):
    """A reader takes its share of the cores: all of them alone, 1/N under torchrun with N ranks on the host
    (LOCAL_WORLD_SIZE), 1/files when session.prefetch announces that many concurrent reads, the whole host again when one
    rank reads for everybody; three quarters of the share inflate, a third parses; the knobs override."""
    import ctypes
    import os
    from epilogos_b200 import _lib

    def budget():
        a, b = ctypes.c_int32(0), ctypes.c_int32(0)
        _lib.call("epi_reader_threads", ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value
    for name in ("EPI_INFLATE_THREADS", "EPI_PARSE_THREADS", "LOCAL_WORLD_SIZE"):
        monkeypatch.delenv(name, raising=False)
    cores = len(os.sched_getaffinity(0))
    alone = budget()
    want_inflate = min(12, (cores * 3 + 2) // 4)
    assert alone == (want_inflate if want_inflate >= 2 else 1, max(1, min(8, (cores + 2) // 3)))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", str(cores))               # one core per rank: the sequential decoder, one parser
    assert budget() == (1, 1)
    _lib.call("epi_reader_concurrency", 1, 1)                        # ... unless one rank reads for everybody
    try:
        assert budget() == alone
    finally:
        _lib.call("epi_reader_concurrency", 0, 0)
    monkeypatch.delenv("LOCAL_WORLD_SIZE")
    _lib.call("epi_reader_concurrency", cores, 0)                    # as many files at once as there are cores
    try:
        assert budget() == (1, 1)
    finally:
        _lib.call("epi_reader_concurrency", 0, 0)
    monkeypatch.setenv("EPI_INFLATE_THREADS", "5")
    monkeypatch.setenv("EPI_PARSE_THREADS", "7")
    assert budget() == (5, 7)


def test_gzip_trailer_crc_is_checked_by_both_crc_paths(tmp_path, monkeypatch):
    """The CRC-32 of the trailer check runs on carry-less multiplies where the CPU has them (csrc/crc_clmul.cpp) and on zlib's
    tables otherwise and for short pieces: members of every length around the 64-byte folding width and the 256-byte
    switch-over decode, and a flipped CRC byte is still caught -- sequential and parallel decoder."""
    rng = np.random.default_rng(31)
    data = rng.integers(32, 127, 3_000_000, dtype=np.uint8).tobytes()
    sizes = [0, 1, 15, 16, 17, 63, 64, 65, 127, 128, 255, 256, 257, 319, 320, 1000, 4095, 65536, 65537, 700001]
    blob, at = b"", 0
    for n in sizes:
        blob += gzip.compress(data[at:at + n], 6)
        at += n
    p = tmp_path / "members.gz"
    p.write_bytes(blob)
    for threads, chunk in ((1, None), (3, 60000)):
        monkeypatch.setenv("EPI_INFLATE_THREADS", str(threads))
        if chunk:
            monkeypatch.setenv("EPI_INFLATE_CHUNK", str(chunk))
        assert _inflate(p, monkeypatch) == data[:at]
        bad = bytearray(blob)
        bad[-6] ^= 0x20                                              # CRC-32 field of the last (700,001-byte) member
        p.write_bytes(bytes(bad))
        with pytest.raises(EpilogosB200Error):
            _inflate(p, monkeypatch)
        p.write_bytes(blob)


def test_simd_and_scalar_row_parsers_agree(tmp_path, monkeypatch):
    """The 16-bytes-at-a-time label tokenizer (csrc/label_simd.cpp, rows of at least 48 columns) against the scalar loops
    (EPI_PARSE_SCALAR=1): same labels for every mix of one-, two- and three-digit labels and every row width around the
    window and tail boundaries, and the same error -- row, column and reason -- for every kind of damaged row."""
    rng = np.random.default_rng(21)

    def both(path, k):
        monkeypatch.delenv("EPI_PARSE_SCALAR", raising=False)
        try:
            a = helpers.read_matrix(path, num_states=k)[1].copy()
        except EpilogosB200Error as exc:
            a = str(exc)
        monkeypatch.setenv("EPI_PARSE_SCALAR", "1")
        try:
            b = helpers.read_matrix(path, num_states=k)[1].copy()
        except EpilogosB200Error as exc:
            b = str(exc)
        monkeypatch.delenv("EPI_PARSE_SCALAR", raising=False)
        return a, b

    for cols in (47, 48, 49, 63, 64, 65, 66, 79, 80, 81, 100, 127, 300, 833):
        for k in (5, 9, 10, 18, 99, 100, 127):
            x = rng.integers(1, k + 1, (40, cols))
            x[rng.random(x.shape) < 0.5] = k
            p = tmp_path / "m.txt"
            p.write_text("".join("chr%d\t%d\t%d\t%s\n" % (i % 3, i * 200, i * 200 + 200, "\t".join(map(str, r)))
                                 for i, r in enumerate(x.tolist())))
            a, b = both(p, k)
            assert isinstance(a, np.ndarray) and np.array_equal(a, x - 1) and np.array_equal(b, x - 1), (cols, k)
    p.write_text("".join("chr1\t%d\t%d\t%s\r\n" % (i * 200, i * 200 + 200, "\t".join(map(str, r))) for i, r in enumerate(x.tolist())))
    a, b = both(p, 127)                                              # CRLF line ends
    assert np.array_equal(a, x - 1) and np.array_equal(b, x - 1)
    # damaged rows: whatever the tokenizer meets, the error is the scalar parser's
    x = rng.integers(1, 19, (30, 200))
    rows = ["chr1\t%d\t%d\t%s" % (i * 200, i * 200 + 200, "\t".join(map(str, r))) for i, r in enumerate(x.tolist())]
    for trial in range(160):
        broken = list(rows)
        r = int(rng.integers(0, len(rows)))
        f = broken[r].split("\t")
        c = 3 + int(rng.integers(0, 200))
        kind = trial % 8
        if kind == 0:
            f[c] = "19"
        elif kind == 1:
            f[c] = "0"
        elif kind == 2:
            f[c] = ""
        elif kind == 3:
            f[c] = "1x"
        elif kind == 4:
            f[c] = "-3"
        elif kind == 5:
            f[c] = "007"
        elif kind == 6:
            f = f[:c] + f[c + 1:]                                    # a column short
        else:
            f = f[:c] + ["7"] + f[c:]                                # a column too many
        broken[r] = "\t".join(f)
        p.write_text("\n".join(broken) + "\n")
        a, b = both(p, 18)
        if kind == 5:                                                # leading zeros are an integer all the same (as for pandas)
            want = x - 1
            want[r, c - 3] = 6
            assert isinstance(a, np.ndarray) and np.array_equal(a, want) and np.array_equal(b, want), trial
            continue
        # (the column count is taken from the first row: a first row with a column more or less shifts the blame to row 1)
        blamed = 1 if (r == 0 and kind in (6, 7)) else r
        assert isinstance(a, str) and a == b and ("row %d" % blamed) in a, (trial, kind, a, b)


def test_native_inflate_reads_the_writers_multi_member_files_across_blocks(tmp_path, monkeypatch):
    """What the score writer emits (one gzip member per 16384 rows) read back by the native decoder: dozens of member
    boundaries, some of them inside a 16 MB reader block, CRC-32 / length checked for each; and the score reader on top."""
    rng = np.random.default_rng(9)
    rows, k = 600_000, 6
    sc = (rng.gamma(0.5, 1.0, (rows, k)) * (rng.random((rows, 1)) < 0.4)).astype(np.float32)
    loc = dict(chrom=np.array(["chr1"] * rows, dtype=object), start=np.arange(rows, dtype=np.int64) * 200,
               end=np.arange(rows, dtype=np.int64) * 200 + 200)
    p = tmp_path / "scores_big.txt.gz"
    writer.write_scores_text(p, sc, loc)
    with gzip.open(p, "rb") as f:
        text = f.read()
    assert len(text) > 2 * (16 << 20)
    assert _inflate(p, monkeypatch) == text
    assert _inflate(p, monkeypatch, use_zlib=True) == text
    monkeypatch.delenv("EPI_ZLIB_INFLATE", raising=False)
    _, got = helpers.read_scores(p)
    assert got.shape == (rows, k) and np.array_equal(got[::997], np.array([[float("%.5f" % float(v)) for v in r] for r in sc[::997]]))


# ---- one gzip stream decoded by several threads (csrc/parallel_inflate.h) --------------------------------------------
def _reader_stats():
    import ctypes
    from epilogos_b200 import _lib
    out = (ctypes.c_int64 * 4)()
    _lib.call("epi_reader_stats", out)
    return list(out)            # mode (2 = parallel), chunks, chunks with a start, chunks accepted


def _matrix_text(rng, rows, cols, k=18, dominant=0.6):
    x = rng.integers(1, k + 1, size=(rows, cols))
    x[rng.random((rows, cols)) < dominant] = k
    return b"".join(b"chr1\t%d\t%d\t" % (i * 200, i * 200 + 200) + b"\t".join(str(int(v)).encode() for v in r) + b"\n"
                    for i, r in enumerate(x))


def _bgzf(data, block=60000, level=6):
    """bgzip's container: independent members of at most 64 KiB with an FEXTRA 'BC' subfield holding the member size."""
    import struct
    import zlib
    out = b""
    for i in range(0, max(len(data), 1), block):
        piece = data[i:i + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = co.compress(piece) + co.flush()
        bsize = 12 + 6 + len(raw) + 8 - 1
        out += b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + raw + \
            struct.pack("<II", zlib.crc32(piece), len(piece) & 0xffffffff)
    return out + b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 27) + \
        b"\x03\0" + b"\0" * 8                                                     # bgzip's empty end-of-file member


def test_parallel_inflate_equals_the_data_for_every_stream_shape(tmp_path, monkeypatch):
    """Chunks of the compressed file decoded concurrently from block / member starts found by inspection, with unknown
    history as markers, chained and resolved in order: the text must be the data for label matrices, low-entropy and
    literal-only streams, runs with short periods, stored and fixed blocks in between, flush points, multi-member and bgzip
    files, for chunks much smaller than, comparable to and larger than a DEFLATE block, and the parallel path must
    actually have been taken for the well-conditioned ones."""
    import zlib
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"0123456789\t\n", dtype=np.uint8)
    matrix = _matrix_text(rng, 12000, 60)
    cases = {
        "matrix": matrix,
        "lowentropy": rng.choice(alphabet, 1_500_000, p=[.3, .2, .1, .05, .05, .05, .05, .04, .03, .03, .08, .02]).tobytes(),
        "random": rng.integers(0, 256, 400000, dtype=np.uint8).tobytes(),
        "runs": b"a" * 70000 + b"ab" * 40000 + b"abc" * 30000 + b"abcdefg" * 20000 + bytes(range(256)) * 3000 + matrix[:300000],
        "mixed": matrix[:400000] + rng.integers(0, 256, 150000, dtype=np.uint8).tobytes() + b"\0" * 300000 + matrix[400000:900000],
    }
    taken = {}
    for name, data in cases.items():
        variants = {"l1": gzip.compress(data, 1), "l6": gzip.compress(data, 6), "l9": gzip.compress(data, 9)}
        step = len(data) // 3 + 1
        variants["members3"] = b"".join(gzip.compress(data[i:i + step], 6) for i in range(0, len(data), step))
        variants["bgzf"] = _bgzf(data)
        co = zlib.compressobj(6, zlib.DEFLATED, 31)
        out = b""
        for j, i in enumerate(range(0, len(data), 50021)):
            out += co.compress(data[i:i + 50021]) + co.flush(zlib.Z_SYNC_FLUSH if j % 2 else zlib.Z_FULL_FLUSH)
        variants["flush"] = out + co.flush() + b"\0" * 19
        co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
        variants["fixed"] = co.compress(data) + co.flush()
        for tag, packed in variants.items():
            p = tmp_path / ("%s_%s.gz" % (name, tag))
            p.write_bytes(packed)
            for chunk, threads in ((1500, 3), (20000, 2), (20000, 5), (150000, 4)):
                monkeypatch.setenv("EPI_INFLATE_CHUNK", str(chunk))
                monkeypatch.setenv("EPI_INFLATE_THREADS", str(threads))
                assert _inflate(p, monkeypatch) == data, (name, tag, chunk, threads)
                st = _reader_stats()
                assert st[0] in (1, 2), (name, tag, chunk, threads, st)          # never the zlib fall-back on valid input
                taken[(name, tag, chunk)] = st
            p.unlink()
    # dynamic-Huffman streams cut into chunks of a few blocks (the first chunks are smaller): most chunks have a start
    # and every start that was found is accepted
    for name in ("matrix", "lowentropy", "mixed"):
        for tag in ("l6", "l9", "members3", "bgzf"):
            st = taken[(name, tag, 150000)]
            assert st[0] == 2 and st[2] >= 3 and st[2] * 2 >= st[1] and st[3] == st[2], (name, tag, st)
    assert taken[("matrix", "fixed", 20000)][0] == 1                             # fixed blocks only: nothing to start from
    monkeypatch.setenv("EPI_INFLATE_THREADS", "1")                               # one thread: the sequential decoder
    p = tmp_path / "seq.gz"
    p.write_bytes(gzip.compress(matrix, 6))
    assert _inflate(p, monkeypatch) == matrix and _reader_stats()[0] == 1


def test_parallel_inflate_is_not_fooled_by_embedded_streams(tmp_path, monkeypatch):
    """The finder looks at raw bytes, so a stored block that holds another gzip file makes it find member and block starts
    that are not starts of THIS stream; a chunk is only accepted where the previous accepted chunk ended, so the text is
    still the data."""
    rng = np.random.default_rng(12)
    inner = gzip.compress(_matrix_text(rng, 6000, 40), 6)
    blob = b"".join(rng.integers(0, 256, 3000, dtype=np.uint8).tobytes() + inner for _ in range(4))
    for level in (0, 6):                                            # stored outright / stored because incompressible
        p = tmp_path / ("outer%d.gz" % level)
        p.write_bytes(gzip.compress(blob, level))
        for chunk in (5000, 40000):
            monkeypatch.setenv("EPI_INFLATE_CHUNK", str(chunk))
            monkeypatch.setenv("EPI_INFLATE_THREADS", "4")
            assert _inflate(p, monkeypatch) == blob, (level, chunk)
    # a real stream followed by stored blocks with embedded streams: parallel mode is taken and false starts are dropped
    data = _matrix_text(rng, 15000, 40) + blob + _matrix_text(rng, 8000, 40)
    p = tmp_path / "both.gz"
    p.write_bytes(gzip.compress(data, 6))
    monkeypatch.setenv("EPI_INFLATE_CHUNK", "30000")
    assert _inflate(p, monkeypatch) == data
    st = _reader_stats()
    assert st[0] == 2 and st[3] < st[2], st                         # some chunk starts were found and NOT accepted


def test_parallel_inflate_rejects_damaged_streams(tmp_path, monkeypatch):
    """Cut, CRC, length and data damage, garbage after the last member and seeded single-bit flips with the parallel
    decoder switched on for a small file: an error or the exact text, never different text."""
    rng = np.random.default_rng(13)
    text = _matrix_text(rng, 9000, 40)
    good = gzip.compress(text, 6)
    assert len(good) > 100000
    monkeypatch.setenv("EPI_INFLATE_CHUNK", "30000")
    monkeypatch.setenv("EPI_INFLATE_THREADS", "3")
    p = tmp_path / "m.txt.gz"
    p.write_bytes(good)
    assert _inflate(p, monkeypatch) == text and _reader_stats()[0] == 2
    loc, states = helpers.read_matrix(p, num_states=18)
    assert states.shape == (9000, 40)
    third = len(good) // 3
    damaged = {"cut in the data": good[:len(good) // 2], "cut in the trailer": good[:-3],
               "crc": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:], "length": good[:-1] + bytes([good[-1] ^ 0x40]),
               "data": good[:third] + bytes([good[third] ^ 0x10]) + good[third + 1:],
               "garbage": good + b"this is not gzip data, 1234567890 1234567890"}
    for what, blob in damaged.items():
        p.write_bytes(blob)
        with pytest.raises(EpilogosB200Error):
            _inflate(p, monkeypatch)
        with pytest.raises(EpilogosB200Error):
            helpers.read_matrix(p, num_states=18)
    p.write_bytes(good + b"\0" * 100 + b"")
    assert _inflate(p, monkeypatch) == text                        # zero padding is tolerated, as by Python's gzip module
    p.write_bytes(good + good)
    assert _inflate(p, monkeypatch) == text + text
    for trial in range(200):
        blob = bytearray(good)
        i = int(rng.integers(2, len(blob)))
        blob[i] ^= 1 << int(rng.integers(0, 8))
        p.write_bytes(bytes(blob))
        try:
            got = _inflate(p, monkeypatch)
        except EpilogosB200Error:
            continue
        assert got == text, "byte %d flipped: accepted with different content" % i


def test_parallel_inflate_hands_over_when_a_chunk_expands_too_much(tmp_path, monkeypatch):
    """A constant matrix compresses ~500-fold: a chunk of the parallel decoder would hold hundreds of megabytes of symbols.
    Single-member files are recognised up front (ISIZE against the file size) and stay with the sequential decoder; behind a
    second member the size is not visible, the chunk hits its budget and zlib finishes the file.  Same text either way."""
    row = b"chr1\t0\t200\t" + b"\t".join([b"18"] * 833) + b"\n"
    data = row * 56000                                               # 140 MB of text in ~270 KB
    tail = b"chr1\t0\t200\t1\t2\n" * 1000
    monkeypatch.setenv("EPI_INFLATE_CHUNK", "500000")               # chunks of 62.5, 62.5 and 125 KB: the third expands to 69 MB
    monkeypatch.setenv("EPI_INFLATE_THREADS", "2")
    p = tmp_path / "const.gz"
    p.write_bytes(gzip.compress(data, 6))
    assert _inflate(p, monkeypatch) == data and _reader_stats()[0] == 1          # never started
    p.write_bytes(gzip.compress(data, 6) + gzip.compress(tail, 6))
    assert _inflate(p, monkeypatch) == data + tail and _reader_stats()[0] == 3   # started, gave up, zlib finished


def test_parallel_and_sequential_readers_parse_the_same_matrix(tmp_path, monkeypatch):
    """read_matrix over a file large enough for the parallel decoder by default (no knobs): same labels and coordinates as
    with one inflate thread, and as the matrix that was written."""
    rng = np.random.default_rng(14)
    x = rng.integers(0, 18, size=(60000, 300)).astype(np.int8)
    x[rng.random(x.shape) < 0.5] = 17
    p = tmp_path / "epilogos_matrix_chr1.txt.gz"
    sys.path.insert(0, str(ROOT))
    from oracle import reference_driver as ref
    ref.write_matrix_tsv_gz(p, x, level=6)
    assert p.stat().st_size > (2 << 20)
    monkeypatch.delenv("EPI_INFLATE_CHUNK", raising=False)
    monkeypatch.setenv("EPI_INFLATE_THREADS", "4")
    loc_p, got_p = helpers.read_matrix(p, num_states=18)
    assert _reader_stats()[0] == 2
    monkeypatch.setenv("EPI_INFLATE_THREADS", "1")
    loc_s, got_s = helpers.read_matrix(p, num_states=18)
    assert _reader_stats()[0] == 1
    assert np.array_equal(got_p, x) and np.array_equal(got_s, x)
    assert np.array_equal(loc_p["start"], loc_s["start"]) and np.array_equal(loc_p["end"], np.arange(1, 60001) * 200)


def _pack_reference(x, cols, bits):
    """Label j of a group of 8 in bits [bits*j, bits*(j+1)) of the little-endian group; rows padded to 16 bytes."""
    bins = x.shape[0]
    groups = (cols + 7) // 8
    lab = np.zeros((bins, groups * 8), dtype=np.uint64)
    lab[:, :cols] = x[:, :cols].astype(np.uint8)
    v = np.zeros((bins, groups), dtype=np.uint64)
    for j in range(8):
        v |= lab[:, j::8] << np.uint64(bits * j)
    out = np.zeros((bins, (groups * bits + 15) // 16 * 16), dtype=np.uint8)
    for b in range(bits):
        out[:, b:groups * bits:bits] = ((v >> np.uint64(8 * b)) & np.uint64(255)).astype(np.uint8)
    return out


@pytest.mark.parametrize("cols,k", [(833, 18), (127, 15), (8, 16), (1, 2), (17, 32), (1000, 25)])
def test_host_packer_layout(cols, k):
    """epi_pack_states_host (CPU threads, no GPU): the bit-packed transport layout of the state matrix, 4 bits per label
    for <= 16 states, else 5, against a numpy restatement of the layout; pad bytes of the int8 rows are ignored."""
    from epilogos_b200 import engine
    rng = np.random.default_rng(cols * 100 + k)
    bins = 531
    x = np.full((bins, engine.pitch_for(cols)), 77, dtype=np.int8)           # garbage in the pad columns
    x[:, :cols] = rng.integers(0, k, size=(bins, cols))
    packed, bits = engine.pack_bits_host(x, cols, k, threads=3)
    assert bits == (4 if k <= 16 else 5)
    assert packed.shape == (bins, engine.packed_pitch(cols, bits))
    assert np.array_equal(packed.numpy(), _pack_reference(x, cols, bits))


def test_writer_non_finite_values_print_like_python(tmp_path):
    """scores.writeScores formats with Python's "{:.5f}": nan (never "-nan"), inf, -inf."""
    vals = np.array([[np.nan, -np.nan, np.inf, -np.inf, 1.5]], dtype=np.float32)
    vals.view(np.uint32)[0, 1] |= 0x80000000                        # a NaN with the sign bit set
    loc = dict(chrom=np.array(["chr1"], dtype=object), start=np.zeros(1, np.int64), end=np.full(1, 200, np.int64))
    path = tmp_path / "nf.txt.gz"
    writer.write_scores_text(path, vals, loc)
    got = gzip.open(path, "rb").read().decode()
    assert got == "chr1\t0\t200\t" + "\t".join("{:.5f}".format(float(v)) for v in vals[0]) + "\n"
    assert got.split("\t")[3:5] == ["nan", "nan"]


def test_trailing_garbage_after_gzip_members_is_an_error(tmp_path):
    """Python's gzip module (the reference's reader) raises on bytes after the last member that are not another member; zero
    padding is tolerated.  zlib's gzread would silently accept both."""
    rows = "".join("chr1\t%d\t%d\t1\t2\t3\n" % (i * 200, i * 200 + 200) for i in range(50))
    good = gzip.compress(rows.encode())
    ok = tmp_path / "ok.txt.gz"
    ok.write_bytes(good + b"\0" * 64)
    assert helpers.read_matrix(ok, num_states=3)[1].shape == (50, 3)
    two = tmp_path / "two.txt.gz"
    two.write_bytes(good + good)
    assert helpers.read_matrix(two, num_states=3)[1].shape == (100, 3)
    bad = tmp_path / "bad.txt.gz"
    bad.write_bytes(good + b"this is not gzip data, 1234567890 1234567890")
    with pytest.raises(EpilogosB200Error):
        helpers.read_matrix(bad, num_states=3)
    with pytest.raises(Exception):
        gzip.open(bad, "rb").read()                                  # the reference's reader refuses it as well
