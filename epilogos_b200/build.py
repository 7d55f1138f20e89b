"""Build the sm_100a shared library in-tree:  epilogos_b200/libepilogos_b200.so

    python -m epilogos_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the gpurun
snapshot, so nothing is compiled there.  cudart is linked statically: the library has no load-time
dependency on libcuda / libcudart and can be dlopen()ed (and its symbol table checked) on a CPU-only host.
"""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libepilogos_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-O3,-Wall", "-cudart", "static"]


def sources():
    """CUDA sources and the plain C++ ones (host code that wants the host compiler alone, e.g. SIMD intrinsics behind a
    target attribute); nvcc hands a .cpp straight to g++."""
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def _stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "epilogos_b200.h",
                                                                  Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objdir = PKG / "csrc" / "_obj"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src.name, out))
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError("epilogos_b200: nvcc compilation failed")
    link = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)] + \
        [str(o) for o in objs] + ["-lz"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("epilogos_b200: link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
