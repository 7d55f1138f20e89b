"""`epilogos` command line mirror (epilogos_b200.run): flag validation on CPU, full local pipeline on the GPU."""
import gzip

import numpy as np
import pytest
from click.testing import CliRunner

from test_host_stages import write_tsv

META = "zero_index\tone_index\tshort_name\tlong_name\n" + "".join("%d\t%d\tS%d\tstate %d\n" % (i, i + 1, i + 1, i + 1)
                                                                      for i in range(18))


def test_flag_validation_messages(tmp_path):
    from epilogos_b200 import run
    r = CliRunner().invoke(run.main, ["-l", "-o", str(tmp_path), "-j", "x"])
    assert "ERROR: [-i, --input-directory] required in 'single' group mode" in r.output
    r = CliRunner().invoke(run.main, ["-m", "paired", "-a", str(tmp_path), "-o", str(tmp_path), "-j", "x"])
    assert "ERROR: [-b, --directory-two] required in 'paired' group mode" in r.output
    r = CliRunner().invoke(run.main, ["-i", str(tmp_path), "-o", str(tmp_path), "-j", "x", "-n"])
    assert "not compatible with [-n, --null-distribution]" in r.output
    meta = tmp_path / "meta.tsv"
    meta.write_text(META)
    assert run.getNumStates(meta) == 18 and run.getStateNames(meta)[17] == "S18"
    inp = tmp_path / "in"
    inp.mkdir()
    (inp / "f.txt").write_text("chr1\t0\t200\t1\n")
    r = CliRunner().invoke(run.main, ["-i", str(inp), "-o", str(tmp_path / "o"), "-j", str(meta), "-s", "4"])
    assert isinstance(r.exception, ValueError) and "Saliency Metric Invalid" in str(r.exception)
    r = CliRunner().invoke(run.main, ["-i", str(tmp_path / "missing"), "-o", str(tmp_path / "o"), "-j", str(meta)])
    assert isinstance(r.exception, FileNotFoundError)
    r = CliRunner().invoke(run.main, ["-v"])
    assert "Version:" in r.output


@pytest.mark.gpu
def test_local_single_pipeline_end_to_end(tmp_path, golden):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from epilogos_b200 import run, session
    from oracle import epilogos_oracle as orc, roi_oracle
    session.clear()
    g = golden("real10_chr1_k18")
    x = g["x"]
    inp = tmp_path / "mydata"; out = tmp_path / "out"
    inp.mkdir()
    write_tsv(inp / "epilogos_matrix_chr1.txt.gz", x, gz=True)
    meta = tmp_path / "meta.tsv"
    meta.write_text(META)
    r = CliRunner().invoke(run.main, ["-l", "-i", str(inp), "-o", str(out), "-j", str(meta), "-s", "1", "-w", "20"])
    assert r.exit_code == 0, r.output + repr(r.exception)
    with gzip.open(out / "scores_mydata_s1_epilogos_matrix_chr1.txt.gz", "rb") as f:
        text = f.read()
    ref_lines = g["s1_text"].tobytes().split(b"\n")
    got_lines = text.split(b"\n")
    assert len(ref_lines) == len(got_lines) and sum(a != b for a, b in zip(ref_lines, got_lines)) <= 2
    # step 4 consumed the temp files and wrote the ROI list; compare with the oracle on the reference's scores
    assert not list(out.glob("temp_scores_*.npz")) and not (out / "exp_freq_mydata_s1.npy").exists()
    starts = np.arange(len(x), dtype=np.int64) * 200
    sel = roi_oracle.max_mean(starts, starts + 200, g["s1_scores"].sum(axis=1), 20, 100)
    states = roi_oracle.max_states(g["s1_scores"], sel["original_idx"], 20)
    want = roi_oracle.roi_text(np.array(["chr1"] * len(x), dtype=object), sel, states, ["S%d" % i for i in range(1, 19)])
    got = (out / "regionsOfInterest_mydata_s1.txt").read_text()
    assert [l.split("\t")[:4] for l in got.splitlines()] == [l.split("\t")[:4] for l in want.splitlines()]


def test_cli_main_end_to_end_on_cpu_with_the_oracle_backend(tmp_path, golden, monkeypatch):
    """The whole `epilogos -l` command path (argument checks, file discovery, prefetch, steps 1-4, clean-up) without a
    GPU: the session's backend is replaced by the oracle-backed stand-in of the host-stage tests, so this covers
    run.main / run.run_stages themselves; the GPU test above covers the same path with the CUDA backend."""
    from fake_backend import OracleBackend
    from epilogos_b200 import run, session
    from oracle import epilogos_oracle as orc
    session.clear()
    monkeypatch.setattr(session, "_backend", OracleBackend())
    x = golden("real10_chr1_k18")["x"][:1500]
    inp = tmp_path / "mydata"; out = tmp_path / "out"
    inp.mkdir()
    write_tsv(inp / "epilogos_matrix_chr1.txt.gz", x[:900], gz=True)
    write_tsv(inp / "epilogos_matrix_chr2.txt.gz", x[900:], chrom="chr2", gz=True)
    meta = tmp_path / "meta.tsv"
    meta.write_text(META)
    r = CliRunner().invoke(run.main, ["-l", "-i", str(inp), "-o", str(out), "-j", str(meta), "-s", "2", "-w", "20"])
    assert r.exit_code == 0, r.output + repr(r.exception)
    assert "STEP 1" in r.output and "STEP 4" in r.output
    exp = orc.normalize_expected(orc.s2_expected_counts(x, 18))
    for name, part, chrom in (("chr1", x[:900], "chr1"), ("chr2", x[900:], "chr2")):
        with gzip.open(out / ("scores_mydata_s2_epilogos_matrix_%s.txt.gz" % name), "rb") as f:
            text = f.read()
        starts = np.arange(len(part)) * 200
        assert text == orc.format_scores_text(orc.s2_scores(part, 18, exp), chrom, starts, starts + 200)
    assert (out / "regionsOfInterest_mydata_s2.txt").exists()
    assert not list(out.glob("temp_*")) and not (out / "exp_freq_mydata_s2.npy").exists()
    session.clear()
    # the score -> ROI hand-over happened in memory (no temp_scores npz was ever written) and gives the very file that
    # the reference's route through temp_scores_*.npz gives (EPILOGOS_B200_KEEP_TEMP=1)
    assert not session.handover
    from epilogos_b200 import helpers
    written = []
    real = helpers.savez_level
    monkeypatch.setattr(helpers, "savez_level", lambda path, *a, **k: (written.append(str(path)), real(path, *a, **k))[1])
    out1 = tmp_path / "out_mem"
    r = CliRunner().invoke(run.main, ["-l", "-i", str(inp), "-o", str(out1), "-j", str(meta), "-s", "2", "-w", "20"])
    assert r.exit_code == 0 and not [w for w in written if "temp_scores" in w]
    monkeypatch.setenv("EPILOGOS_B200_KEEP_TEMP", "1")
    session.clear()
    out2 = tmp_path / "out_files"
    r = CliRunner().invoke(run.main, ["-l", "-i", str(inp), "-o", str(out2), "-j", str(meta), "-s", "2", "-w", "20"])
    assert r.exit_code == 0 and len([w for w in written if "temp_scores" in w]) == 2
    assert (out1 / "regionsOfInterest_mydata_s2.txt").read_text() == (out2 / "regionsOfInterest_mydata_s2.txt").read_text()
    assert (out / "regionsOfInterest_mydata_s2.txt").read_text() == (out2 / "regionsOfInterest_mydata_s2.txt").read_text()
    session.clear()


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference`: rank 0 prints ONE JSON line with the contract's keys (the unmodified reference staged under
    oracle/_ref when /root/reference exists, else the oracle port), the same `config` object our arm prints, e2e = value;
    every other rank exits 0 without output."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, EPI_BENCH_REF_BUDGET_S="3")
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--config", "s2_genome_127"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "bins/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["biosamples"] == 127 and d["config"]["states"] == 15 and "workload" in d["config"]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""
