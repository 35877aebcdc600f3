"""In-tree nvcc build of the stminer_b200 C-ABI library (sm_100a only).

``python -m stminer_b200.build`` (or ``__graft_entry__.build()``) compiles every ``csrc/*.cu`` into
``stminer_b200/csrc/libstminer_b200.so``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libstminer_b200.so")
BUILD_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the stminer_b200 CUDA library cannot be built")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _source_hash() -> str:
    """Content hash of everything the library is compiled from (mtimes do not survive a snapshot copy)."""
    import hashlib
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    deps.append(os.path.join(HERE, "..", "include", "stminer_b200.h"))
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    """True when the library is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(LIB_PATH + ".srchash"):
        return True
    with open(LIB_PATH + ".srchash") as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False, defines=None, out: str = None) -> str:
    """Compile csrc/*.cu into the in-tree library.  ``defines`` / ``out`` build a tuning variant
    (e.g. ``defines=["STM_EM_MINB_SMALL=5"]``) next to the default library."""
    lib_path = out or LIB_PATH
    if not force and not defines and not out and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    build_dir = BUILD_DIR if not out else BUILD_DIR + "_" + os.path.basename(out)
    os.makedirs(build_dir, exist_ok=True)
    objs = []
    dflags = ["-D" + d for d in (defines or [])]

    def compile_one(src):
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *dflags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if lib_path == LIB_PATH and not defines:
        with open(LIB_PATH + ".srchash", "w") as f:
            f.write(_source_hash())
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
