"""Stage the UNMODIFIED reference package next to the oracle so that it can travel to the GPU box.

TEST / BENCHMARK INFRASTRUCTURE ONLY.  The reference (meuleman/epilogos) is pure Python, so there is nothing to
compile: "building" oracle/_ref means copying the package's own source files, byte for byte, from where they lie under
/root/reference into oracle/_ref/epilogos/.  oracle/_ref/ is git-ignored (the sources never enter this repository's
history) but NOT gpurun-ignored, so the staged copy travels with the snapshot like a built .so and `bench.py --impl
reference` / `cpu_baseline` can time the reference's real `expected.main -> expectedCombination.main -> scores.main`
path (TSV.gz parse and gz write included, run.py:191-303) on the GPU box's own host cores.

    python -m oracle.stage_reference          # called by __graft_entry__.build() when /root/reference exists

A MANIFEST with the sha256 of every staged file is written so that a run can state exactly what it timed.
"""
import hashlib
import json
import shutil
from pathlib import Path

SOURCE = Path("/root/reference")
DEST = Path(__file__).resolve().parent / "_ref"

# the modules the scoring path imports (expected / expectedCombination / scores -> helpers -> filter_regions)
FILES = ["epilogos/__init__.py", "epilogos/expected.py", "epilogos/expectedCombination.py", "epilogos/scores.py",
         "epilogos/helpers.py", "epilogos/filter_regions.py"]


def staged_root():
    """Directory to put on sys.path to import the reference: /root/reference when it exists (authoring container),
    else the staged copy, else None."""
    if (SOURCE / "epilogos" / "scores.py").is_file():
        return SOURCE
    if (DEST / "epilogos" / "scores.py").is_file():
        return DEST
    return None


def stage(verbose=True):
    if not (SOURCE / "epilogos" / "scores.py").is_file():
        if verbose:
            print("stage_reference: %s not present, nothing staged" % SOURCE)
        return None
    manifest = {}
    for rel in FILES:
        src = SOURCE / rel
        if not src.is_file():
            if rel.endswith("__init__.py"):
                (DEST / rel).parent.mkdir(parents=True, exist_ok=True)
                (DEST / rel).write_bytes(b"")
                manifest[rel] = hashlib.sha256(b"").hexdigest()
                continue
            raise FileNotFoundError(src)
        dst = DEST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    version = ""
    setup = SOURCE / "setup.py"
    if setup.is_file():
        for line in setup.read_text().splitlines():
            if "version" in line and "=" in line:
                version = line.strip().strip(",")
                break
    (DEST / "MANIFEST.json").write_text(json.dumps({"source": str(SOURCE), "version_line": version, "sha256": manifest},
                                                   indent=1))
    if verbose:
        print("stage_reference: %d files -> %s" % (len(manifest), DEST))
    return DEST


if __name__ == "__main__":
    stage()
